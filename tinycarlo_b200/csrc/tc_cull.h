// tc_cull.h — host-side builder of the camera pass's visible-set tables (shared by tc_api.cu and the CPU-only test build).
//
// The reference's camera pass (camera.py:52-110) transforms, clips and projects EVERY laneline node of the map, although a
// frame shows the few edges within max_range of the camera. The block-per-env render kernel instead works on a sub-graph
// chosen by the ground cell the camera stands in. The sub-graphs are built here, once per camera configuration, such that
// the emitted segments are the same as those of the whole graph:
//
//  * A node is VISIBLE (camera.py:92-93) only if its final position p is in front of the camera, within max_range, and
//    projects strictly inside the frame; such a p lies within R = max_range * sqrt(1 + tx^2 + ty^2) of the camera centre
//    (tx = max(cx, W-cx)/fx, ty = max(cy, H-cy)/fy), so within R of it on the ground.
//  * The four clip passes (camera.py:70-86) move a node only along an edge towards a neighbour's (possibly moved)
//    position, and each pass reads the previous pass's state of the direct neighbours. After 4 passes a node's position
//    depends on the initial positions of the nodes within 4 hops and lies in their convex hull (both clip planes are
//    crossed between the two endpoints, so every move is an interpolation; the one exception - a front node closer than
//    1e-7 m to the camera plane makes the near-plane move an extrapolation by at most 1e-7 / |cos(edge, optical axis)| -
//    is covered by `margin`: it would need that node on an edge parallel to the camera plane within 2e-6 rad).
//    Hence a visible node v satisfies dist(v, camera) <= R + reach4(v), reach4(v) = max distance from v to a node within
//    4 hops: the CORE set V0 of a cell = nodes with dist(v, cell rectangle) <= R + margin + reach4(v).
//  * An emitted edge has >= 1 visible endpoint (camera.py:95), i.e. an endpoint in V0; its other endpoint is 1 hop away, and
//    the final positions of both depend on nodes within 4 further hops: the sub-graph is the graph induced on all nodes
//    within 5 hops of V0 (edges in list order, adjacency in list order). Nodes on its rim may end up with other
//    positions than in the whole graph, which is harmless: only core nodes may report "visible".
//
// Outside the grid (camera farther than R + margin + max reach4 from every node) the core set is empty: the frame is empty.
// Culling is switched off (one cell = the whole graph) when the camera model does not fit the argument: non-orthonormal
// extrinsics, camera (almost) in the ground plane, non-finite intrinsics, TC_CULL=0, or sub-graphs that are not smaller.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <queue>
#include <vector>

#include "tc_core.cuh"

struct TcCull {
    TcCullGrid grid{};
    std::vector<TcCellBlob> desc;     // nx*ny cells + one empty descriptor (outside the grid); culling off: desc[0] = whole graph
    std::vector<unsigned char> blob;  // the cells' tables, each 16-byte aligned
    int max_nodes = 0, max_edges = 0, max_bytes = 0;
    double radius = -1.0;             // the R the cells were built for (< 0: culling off)
    double mean_nodes = 0.0;          // over the non-empty cells (diagnostics)
};

// appends the tables of the sub-graph induced on `keep` (ascending global node ids) to c.blob; returns its descriptor
static inline TcCellBlob tc_cull_emit(TcCull &c, const double *nodes, const std::vector<int32_t> &edges /*2m*/, const std::vector<uint8_t> &edge_cls,
                                      const std::vector<int32_t> &keep, const std::vector<uint8_t> &core_flag, std::vector<int32_t> &local /*scratch n, -1*/) {
    TcCellBlob d{};
    const int n = (int)keep.size();
    for (int i = 0; i < n; i++) local[keep[i]] = i;
    std::vector<int32_t> ed;
    std::vector<uint8_t> cls;
    const int M = (int)edge_cls.size();
    for (int e = 0; e < M; e++) {
        int a = local[edges[2 * e]], b = local[edges[2 * e + 1]];
        if (a >= 0 && b >= 0) { ed.push_back(a); ed.push_back(b); cls.push_back(edge_cls[e]); }
    }
    const int m = (int)cls.size();
    std::vector<int32_t> oo(n + 1, 0), io(n + 1, 0), oe(std::max(m, 1)), ie(std::max(m, 1));
    for (int e = 0; e < m; e++) { oo[ed[2 * e] + 1]++; io[ed[2 * e + 1] + 1]++; }
    for (int i = 0; i < n; i++) { oo[i + 1] += oo[i]; io[i + 1] += io[i]; }
    {
        std::vector<int32_t> oc(oo.begin(), oo.end() - 1), ic(io.begin(), io.end() - 1);
        for (int e = 0; e < m; e++) { oe[oc[ed[2 * e]]++] = e; ie[ic[ed[2 * e + 1]]++] = e; }
    }
    d.n_nodes = n; d.n_edges = m;
    int32_t o = 0;
    auto sec = [&](int32_t bytes) { int32_t at = o; o = (o + bytes + 15) & ~15; return at; };
    sec(n * 16);
    d.off_edges = sec(m * 8); d.off_out_off = sec((n + 1) * 4); d.off_out_edge = sec(m * 4);
    d.off_in_off = sec((n + 1) * 4); d.off_in_edge = sec(m * 4); d.off_core = sec(n); d.off_edge_cls = sec(m);
    d.bytes = n > 0 ? o : 0;
    d.offset = (int32_t)c.blob.size();
    if (n > 0) {
        c.blob.resize(c.blob.size() + (size_t)o, 0);
        unsigned char *b = c.blob.data() + d.offset;
        for (int i = 0; i < n; i++) {
            memcpy(b + (size_t)i * 16, nodes + 2 * (size_t)keep[i], 16);
            b[d.off_core + i] = core_flag[keep[i]];
        }
        if (m > 0) {
            memcpy(b + d.off_edges, ed.data(), (size_t)m * 8);
            memcpy(b + d.off_out_edge, oe.data(), (size_t)m * 4);
            memcpy(b + d.off_in_edge, ie.data(), (size_t)m * 4);
            memcpy(b + d.off_edge_cls, cls.data(), (size_t)m);
        }
        memcpy(b + d.off_out_off, oo.data(), (size_t)(n + 1) * 4);
        memcpy(b + d.off_in_off, io.data(), (size_t)(n + 1) * 4);
    }
    for (int i = 0; i < n; i++) local[keep[i]] = -1;
    c.max_nodes = std::max(c.max_nodes, n); c.max_edges = std::max(c.max_edges, m); c.max_bytes = std::max(c.max_bytes, (int)d.bytes);
    return d;
}

// radius < 0: culling off. cell: edge length of the ground cells in metres; margin: see the header comment.
static inline void tc_build_cull(const TcMapDesc *map, double radius, double cell, double margin, TcCull &c) {
    c = TcCull();
    const int C = map->n_classes, n = map->ll_node_off[C], M = map->ll_edge_off[C];
    std::vector<int32_t> edges(2 * (size_t)std::max(M, 1));
    std::vector<uint8_t> edge_cls((size_t)M);
    for (int k = 0; k < C; k++)
        for (int e = map->ll_edge_off[k]; e < map->ll_edge_off[k + 1]; e++) {
            edges[2 * e] = map->ll_edges[2 * e] + map->ll_node_off[k];
            edges[2 * e + 1] = map->ll_edges[2 * e + 1] + map->ll_node_off[k];
            edge_cls[e] = (uint8_t)k;
        }
    const double *nodes = map->ll_nodes;
    std::vector<int32_t> local((size_t)std::max(n, 1), -1);
    auto whole_graph = [&]() {
        c = TcCull();
        std::vector<int32_t> keep(n);
        for (int i = 0; i < n; i++) keep[i] = i;
        std::vector<uint8_t> core((size_t)std::max(n, 1), 1);
        c.grid.nx = c.grid.ny = 0;
        c.desc.push_back(tc_cull_emit(c, nodes, edges, edge_cls, keep, core, local));
        c.radius = -1.0;
        c.mean_nodes = n;
    };
    bool finite = std::isfinite(radius) && radius >= 0 && cell > 0 && n > 0;
    for (int i = 0; finite && i < 2 * n; i++) finite = std::isfinite(nodes[i]);
    if (!finite) { whole_graph(); return; }

    // undirected adjacency
    std::vector<std::vector<int32_t>> adj(n);
    for (int e = 0; e < M; e++) { adj[edges[2 * e]].push_back(edges[2 * e + 1]); adj[edges[2 * e + 1]].push_back(edges[2 * e]); }
    // reach4[v]: max distance from v to a node within 4 hops
    std::vector<double> reach4(n, 0.0);
    std::vector<int32_t> depth(n, -1), touched;
    double reach_max = 0.0;
    for (int v = 0; v < n; v++) {
        touched.clear();
        std::queue<int32_t> q;
        depth[v] = 0; q.push(v); touched.push_back(v);
        while (!q.empty()) {
            int u = q.front(); q.pop();
            double dx = nodes[2 * u] - nodes[2 * v], dy = nodes[2 * u + 1] - nodes[2 * v + 1];
            reach4[v] = std::max(reach4[v], std::sqrt(dx * dx + dy * dy));
            if (depth[u] == 4) continue;
            for (int w : adj[u]) if (depth[w] < 0) { depth[w] = depth[u] + 1; q.push(w); touched.push_back(w); }
        }
        for (int u : touched) depth[u] = -1;
        reach_max = std::max(reach_max, reach4[v]);
    }
    double lo[2] = {nodes[0], nodes[1]}, hi[2] = {nodes[0], nodes[1]};
    for (int i = 0; i < n; i++)
        for (int j = 0; j < 2; j++) { lo[j] = std::min(lo[j], nodes[2 * i + j]); hi[j] = std::max(hi[j], nodes[2 * i + j]); }
    const double pad = radius + margin + reach_max + cell;
    c.grid.x0 = lo[0] - pad; c.grid.y0 = lo[1] - pad; c.grid.inv_cell = 1.0 / cell;
    const double fx = std::ceil((hi[0] + pad - c.grid.x0) / cell), fy = std::ceil((hi[1] + pad - c.grid.y0) / cell);
    if (!(fx * fy <= 65536.0)) { whole_graph(); return; }   // the map is huge relative to the cell: not worth the tables
    c.grid.nx = (int32_t)fx; c.grid.ny = (int32_t)fy;
    c.radius = radius;
    std::vector<uint8_t> core(n, 0);
    std::vector<int32_t> keep;
    double sum_nodes = 0.0;
    int non_empty = 0;
    for (int iy = 0; iy < c.grid.ny; iy++)
        for (int ix = 0; ix < c.grid.nx; ix++) {
            // the device computes the cell as floor((x - x0) * inv_cell): widen the rectangle by a rounding guard
            const double guard = 1e-9 * (1.0 + std::fabs(c.grid.x0) + std::fabs(c.grid.y0) + cell * (c.grid.nx + c.grid.ny));
            const double rx0 = c.grid.x0 + ix * cell - guard, rx1 = c.grid.x0 + (ix + 1) * cell + guard;
            const double ry0 = c.grid.y0 + iy * cell - guard, ry1 = c.grid.y0 + (iy + 1) * cell + guard;
            touched.clear();
            std::queue<int32_t> q;
            for (int v = 0; v < n; v++) {
                double dx = std::max(std::max(rx0 - nodes[2 * v], nodes[2 * v] - rx1), 0.0);
                double dy = std::max(std::max(ry0 - nodes[2 * v + 1], nodes[2 * v + 1] - ry1), 0.0);
                if (std::sqrt(dx * dx + dy * dy) <= radius + margin + reach4[v]) { core[v] = 1; depth[v] = 0; q.push(v); touched.push_back(v); }
            }
            while (!q.empty()) {
                int u = q.front(); q.pop();
                if (depth[u] == 5) continue;
                for (int w : adj[u]) if (depth[w] < 0) { depth[w] = depth[u] + 1; q.push(w); touched.push_back(w); }
            }
            keep.assign(touched.begin(), touched.end());
            std::sort(keep.begin(), keep.end());
            c.desc.push_back(tc_cull_emit(c, nodes, edges, edge_cls, keep, core, local));
            if (!keep.empty()) { sum_nodes += keep.size(); non_empty++; }
            for (int u : touched) { depth[u] = -1; core[u] = 0; }
        }
    c.desc.push_back(TcCellBlob{});   // outside the grid: nothing can be visible
    c.mean_nodes = non_empty ? sum_nodes / non_empty : 0.0;
    if (c.max_nodes > 0.85 * n) whole_graph();   // the sub-graphs are not smaller than the map: skip the lookup
}

// The R of one camera row (see the header comment), or < 0 when the row does not fit the argument.
static inline double tc_cull_radius_of(const double *cam /*TC_CAM_N*/, int H, int W) {
    const double *E = cam + TC_CAM_E;
    const double fx = cam[TC_CAM_FX], fy = cam[TC_CAM_FY], cx = cam[TC_CAM_CX], cy = cam[TC_CAM_CY], mr = cam[TC_CAM_MAX_RANGE];
    for (int i = 0; i < TC_CAM_MAX_RANGE + 1; i++) if (!std::isfinite(cam[i])) return -1.0;
    if (!(fx > 0) || !(fy > 0) || !(mr > 0)) return -1.0;
    // rotation part orthonormal?
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += E[4 * k + i] * E[4 * k + j];
            if (std::fabs(s - (i == j ? 1.0 : 0.0)) > 1e-9) return -1.0;
        }
    // camera height above the ground plane (car frame z = world z): centre = -R^T t
    const double cz = -(E[2] * E[3] + E[6] * E[7] + E[10] * E[11]);
    if (std::fabs(cz) < 1e-3) return -1.0;
    const double tx = std::max(cx, (double)W - cx) / fx, ty = std::max(cy, (double)H - cy) / fy;
    if (!(tx >= 0) || !(ty >= 0)) return -1.0;
    return mr * std::sqrt(1.0 + tx * tx + ty * ty);
}

// ------------------------------------------------------------------------------------------------ nearest-laneline index
// car.py:58 asks, per class, for the edge minimising f_e(p) = d(p,n0) + d(p,n1) (layer.py:33-44) - a scan of all edges in the
// reference and, with a float pre-filter, in the tracking kernel. f_e is 2-Lipschitz in p, so for p in a ground cell with
// centre c and half diagonal r the arg-min e* obeys f_e*(c) <= f_e*(p) + 2r <= f_min(p) ... <= f_min(c) + 4r: the candidate
// list of a cell = { e : f_e(c) <= min_e f_e(c) + 4r (+ guard) }, ascending, contains every edge that attains the minimum
// anywhere in the cell, ties included (the first one wins, as in the reference). Outside the grid the kernel scans all edges.
struct TcNear {
    double x0 = 0, y0 = 0, inv_cell = 0;
    int nx = 0, ny = 0;             // 0: no index
    std::vector<int32_t> off;       // [C * nx * ny + 1]
    std::vector<uint16_t> edge;
    double mean_len = 0.0;
    int max_len = 0;
};
static inline void tc_build_near(const TcMapDesc *map, double pad /* metres around the laneline bounding box */, int max_cells, TcNear &nr) {
    nr = TcNear();
    const int C = map->n_classes, n = map->ll_node_off[C];
    if (n <= 0) return;
    const double *nodes = map->ll_nodes;
    for (int i = 0; i < 2 * n; i++) if (!std::isfinite(nodes[i])) return;
    for (int c = 0; c < C; c++) if (map->ll_edge_off[c + 1] - map->ll_edge_off[c] > 65535) return;
    double lo[2] = {nodes[0], nodes[1]}, hi[2] = {nodes[0], nodes[1]};
    for (int i = 0; i < n; i++)
        for (int j = 0; j < 2; j++) { lo[j] = std::min(lo[j], nodes[2 * i + j]); hi[j] = std::max(hi[j], nodes[2 * i + j]); }
    const double w = hi[0] - lo[0] + 2 * pad, h = hi[1] - lo[1] + 2 * pad;
    double cell = std::max(0.02, std::sqrt(w * h / std::max(max_cells, 1)));
    const double fx = std::ceil(w / cell), fy = std::ceil(h / cell);
    if (!(fx >= 1 && fy >= 1 && fx * fy <= 4.0 * max_cells)) return;
    nr.x0 = lo[0] - pad; nr.y0 = lo[1] - pad; nr.inv_cell = 1.0 / cell; nr.nx = (int)fx; nr.ny = (int)fy;
    const double r = cell * 0.70710678118654757 + 1e-9 * (1.0 + std::fabs(nr.x0) + std::fabs(nr.y0) + w + h);   // half diagonal + lookup rounding guard
    const size_t ncell = (size_t)nr.nx * nr.ny;
    nr.off.assign((size_t)C * ncell + 1, 0);
    std::vector<double> f;
    size_t total = 0;
    for (int c = 0; c < C; c++) {
        const double *nd = nodes + 2 * (size_t)map->ll_node_off[c];
        const int32_t *ed = map->ll_edges + 2 * (size_t)map->ll_edge_off[c];
        const int m = map->ll_edge_off[c + 1] - map->ll_edge_off[c];
        f.resize((size_t)std::max(m, 1));
        for (int iy = 0; iy < nr.ny; iy++)
            for (int ix = 0; ix < nr.nx; ix++) {
                const double cx = nr.x0 + (ix + 0.5) * cell, cy = nr.y0 + (iy + 0.5) * cell;
                double fmin = INFINITY;
                for (int e = 0; e < m; e++) {
                    const double *a = nd + 2 * ed[2 * e], *b = nd + 2 * ed[2 * e + 1];
                    f[e] = std::sqrt((cx - a[0]) * (cx - a[0]) + (cy - a[1]) * (cy - a[1])) + std::sqrt((cx - b[0]) * (cx - b[0]) + (cy - b[1]) * (cy - b[1]));
                    fmin = std::min(fmin, f[e]);
                }
                const double lim = fmin + 4.0 * r + 1e-9 * (1.0 + fmin);
                int len = 0;
                for (int e = 0; e < m; e++)
                    if (f[e] <= lim) { nr.edge.push_back((uint16_t)e); len++; }
                nr.off[(size_t)c * ncell + (size_t)iy * nr.nx + ix + 1] = len;
                nr.max_len = std::max(nr.max_len, len);
                total += len;
            }
    }
    for (size_t i = 0; i < (size_t)C * ncell; i++) nr.off[i + 1] += nr.off[i];
    nr.mean_len = (double)total / (double)((size_t)C * ncell);
}
