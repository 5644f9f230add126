"""The reference's own known-answer tests (test/test_layer.py, test/test_helper.py) replayed on the CPU oracle."""
import numpy as np

import layer_kats as K
from oracle import oracle as orc


def one_layer(nodes, edges):
    data = {"lanelines": {"test": {"layer_color": [0, 0, 0], "nodes": [list(map(float, n)) for n in nodes],
                                   "edges": [list(e) for e in edges]}},
            "lanepath": {"layer_color": [0, 0, 0], "nodes": [[0.0, 0.0], [1.0, 0.0]], "edges": [[0, 1]]}}
    return orc.OracleMap(data, 1)


def P(p):
    return np.array(p, np.float64).ctypes.data_as(orc._dp)


def I(a):
    return np.array(a, np.int32).ctypes.data_as(orc._ip)


def test_clip_angle():
    for a, want in K.CLIP_ANGLE:
        assert orc.lib().orc_clip_angle(a) == want


def test_nearest_edge():
    for nodes, edges, cases in K.NEAREST_EDGE:
        m = one_layer(nodes, edges)
        for pos, want in cases:
            assert orc.lib().orc_layer_nearest_edge(m.handle, P(pos)) == want, (pos, want)


def test_nearest_edge_with_orientation():
    for nodes, edges, cases in K.NEAREST_EDGE_ORIENT:
        m = one_layer(nodes, edges)
        for (pos, o), want in cases:
            got = orc.lib().orc_layer_nearest_edge_with_orientation(m.handle, P(pos), o, 30.0)
            assert got == (-1 if want is None else want), (pos, o, want, got)


def test_within_bounds():
    for nodes, edges, cases in K.WITHIN_BOUNDS:
        m = one_layer(nodes, edges)
        for pos, want in cases:
            assert bool(orc.lib().orc_layer_within_bounds(m.handle, P(pos), *edges[0])) == want, (nodes, edges, pos)


def test_distance_to_edge():
    for nodes, edges, cases in K.DISTANCE_TO_EDGE_EXACT:
        m = one_layer(nodes, edges)
        for pos, want in cases:
            assert orc.lib().orc_layer_distance_to_edge(m.handle, P(pos), *edges[0]) == want
    for nodes, edges, cases in K.DISTANCE_TO_EDGE_CLOSE:
        m = one_layer(nodes, edges)
        for pos, want in cases:
            assert abs(orc.lib().orc_layer_distance_to_edge(m.handle, P(pos), *edges[0]) - want) < 1e-5


def test_pick_node():
    for nodes, node, o, conn, want in K.PICK_NODE:
        m = one_layer(nodes, [(0, 1)])
        got = orc.lib().orc_layer_pick_node(m.handle, node, o, I(conn if conn else [0]), len(conn))
        assert got == (-1 if want is None else want), (node, o, conn, want, got)


def test_connected_edge():
    for nodes, edges, pos, edge, o, want in K.CONNECTED_EDGE:
        m = one_layer(nodes, edges)
        out = np.zeros(2, np.int32)
        rc = orc.lib().orc_layer_nearest_connected_edge(m.handle, P(pos), I(edge), o, out.ctypes.data_as(orc._ip))
        if want is None:
            assert rc == -1
        else:
            assert rc == 0 and tuple(out) == want, (pos, edge, want, tuple(out))
