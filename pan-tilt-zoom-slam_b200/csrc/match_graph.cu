// match_graph.cu - match graph -> bundle-adjustment observation list on the device.
//
// Reference behaviour: slam_system/image_process.py:612-650 (build_matching_graph steps 4-5).  The global landmark id of a keypoint
// is decided GREEDILY while the matches are visited in order (pairs (i, j) ascending, matches in list order): a match whose two
// keypoints are both new opens a new id, a match with one labelled end hands that id to the other end, a match between two
// labelled ends changes nothing (it is only reported when the ids differ, :624).  A label never changes once set, therefore
//     label(n) = label(other end of the FIRST match that touches n),  unless that match is also the first match of its other
//     end - then the match opened a new id, numbered in visiting order among all such "root" matches.
// That is a forest of first-match pointers whose chains strictly decrease in match index: first match per keypoint by atomicMin,
// root matches flagged and numbered by a prefix sum, roots found by pointer jumping.  It reproduces the sequential result exactly,
// inconsistent matches included (tests/test_match_graph.py compares with the reference golden and with the host loop on random
// graphs).  The flat observation list follows bundle_adjustment.py:67-98: match k contributes observation 2k (image i, source
// keypoint) and 2k+1 (image j, destination keypoint), both with the landmark id of the SOURCE keypoint.
#include <cub/cub.cuh>

#include "common.h"

namespace {

__global__ void k_mg_first(int n_edge, const int32_t* __restrict__ a, const int32_t* __restrict__ b, int32_t* __restrict__ first) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_edge) return;
    atomicMin(first + a[k], k);
    atomicMin(first + b[k], k);
}

__global__ void k_mg_root(int n_edge, const int32_t* __restrict__ a, const int32_t* __restrict__ b, const int32_t* __restrict__ first,
                          int32_t* __restrict__ root) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_edge) return;
    root[k] = (first[a[k]] == k && first[b[k]] == k) ? 1 : 0;
}

__global__ void k_mg_parent(int n_node, int n_edge, const int32_t* __restrict__ a, const int32_t* __restrict__ b,
                            const int32_t* __restrict__ first, const int32_t* __restrict__ root, int32_t* __restrict__ parent) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_node) return;
    const int e = first[n];
    if (e >= n_edge) { parent[n] = n; return; }                 // keypoint without a match
    parent[n] = root[e] ? n : (a[e] == n ? b[e] : a[e]);
}

__global__ void k_mg_jump(int n_node, int32_t* __restrict__ parent, int* __restrict__ changed) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_node) return;
    const int p = parent[n], pp = parent[p];
    if (pp != p) { parent[n] = pp; *changed = 1; }
}

__global__ void k_mg_label(int n_node, int n_edge, const int32_t* __restrict__ first, const int32_t* __restrict__ parent,
                           const int32_t* __restrict__ root_rank, int32_t* __restrict__ label) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_node) return;
    label[n] = first[n] >= n_edge ? -1 : root_rank[first[parent[n]]];
}

__global__ void k_mg_flatten(int n_edge, const int32_t* __restrict__ a, const int32_t* __restrict__ b, const int32_t* __restrict__ node_img,
                             const double* __restrict__ node_xy, const int32_t* __restrict__ label, int32_t* __restrict__ cam,
                             int32_t* __restrict__ lm, double* __restrict__ xy) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_edge) return;
    const int na = a[k], nb = b[k], l = label[na];
    cam[2 * k] = node_img[na]; cam[2 * k + 1] = node_img[nb];
    lm[2 * k] = l; lm[2 * k + 1] = l;
    const double2 pa = reinterpret_cast<const double2*>(node_xy)[na], pb = reinterpret_cast<const double2*>(node_xy)[nb];
    reinterpret_cast<double2*>(xy)[2 * k] = pa;
    reinterpret_cast<double2*>(xy)[2 * k + 1] = pb;
}

}  // namespace

extern "C" int ptzba_match_graph_to_observations(ptzba_ctx* ctx, int mem, int n_node, const int32_t* node_img, const double* node_xy,
                                                 int n_edge, const int32_t* edge_a, const int32_t* edge_b, int32_t* out_label,
                                                 int32_t* out_n_landmark, int32_t* out_cam, int32_t* out_lm, double* out_xy) {
    if (!ctx) return PTZBA_ERR_ARG;
    ARG_CHECK(ctx, n_node >= 0 && n_edge >= 0 && out_n_landmark && (n_edge == 0 || (edge_a && edge_b)) && (n_node == 0 || out_label));
    ARG_CHECK(ctx, (out_cam == nullptr) == (out_lm == nullptr) && (out_cam == nullptr) == (out_xy == nullptr));
    ARG_CHECK(ctx, out_cam == nullptr || n_node == 0 || (node_img && node_xy));
    *out_n_landmark = 0;
    cudaStream_t s = ctx->stream;
    InArray<int32_t> d_a, d_b, d_img;
    InArray<double> d_xy;
    OutArray<int32_t> o_label, o_cam, o_lm;
    OutArray<double> o_xy;
    DevBuf<int32_t> first, root, rank, parent;
    DevBuf<int> changed;
    DevBuf<unsigned char> tmp;
    CU_CHECK(ctx, d_a.stage(mem, edge_a, (size_t)n_edge, s));
    CU_CHECK(ctx, d_b.stage(mem, edge_b, (size_t)n_edge, s));
    CU_CHECK(ctx, o_label.stage(mem, out_label, (size_t)n_node));
    CU_CHECK(ctx, first.alloc((size_t)n_node + 1)); CU_CHECK(ctx, parent.alloc((size_t)n_node + 1));
    CU_CHECK(ctx, root.alloc((size_t)n_edge + 1)); CU_CHECK(ctx, rank.alloc((size_t)n_edge + 1)); CU_CHECK(ctx, changed.alloc(1));
    if (n_node == 0) return PTZBA_OK;
    // first[n] = n_edge (no match) -> int pattern: fill with a kernel-free trick: 0x7f7f7f7f > any edge index we accept
    ARG_CHECK(ctx, n_edge < 0x7f7f7f7f);
    CU_CHECK(ctx, cudaMemsetAsync(first.p, 0x7f, (size_t)n_node * sizeof(int32_t), s));
    const int T = 256;
    if (n_edge > 0) {
        k_mg_first<<<div_up(n_edge, T), T, 0, s>>>(n_edge, d_a.d, d_b.d, first.p);
        KERNEL_POST(ctx);
        k_mg_root<<<div_up(n_edge, T), T, 0, s>>>(n_edge, d_a.d, d_b.d, first.p, root.p);
        KERNEL_POST(ctx);
        size_t bytes = 0;
        CU_CHECK(ctx, cub::DeviceScan::ExclusiveSum(nullptr, bytes, root.p, rank.p, n_edge, s));
        CU_CHECK(ctx, tmp.alloc(bytes));
        CU_CHECK(ctx, cub::DeviceScan::ExclusiveSum(tmp.p, bytes, root.p, rank.p, n_edge, s));
    }
    k_mg_parent<<<div_up(n_node, T), T, 0, s>>>(n_node, n_edge, d_a.d, d_b.d, first.p, root.p, parent.p);
    KERNEL_POST(ctx);
    for (int it = 0; it < 64; ++it) {                       // pointer jumping: chain lengths halve per round
        int h_changed = 0;
        CU_CHECK(ctx, cudaMemsetAsync(changed.p, 0, sizeof(int), s));
        k_mg_jump<<<div_up(n_node, T), T, 0, s>>>(n_node, parent.p, changed.p);
        KERNEL_POST(ctx);
        CU_CHECK(ctx, cudaMemcpyAsync(&h_changed, changed.p, sizeof(int), cudaMemcpyDeviceToHost, s));
        CU_CHECK(ctx, cudaStreamSynchronize(s));
        if (!h_changed) break;
    }
    k_mg_label<<<div_up(n_node, T), T, 0, s>>>(n_node, n_edge, first.p, parent.p, rank.p, o_label.d);
    KERNEL_POST(ctx);
    int32_t last_rank = 0, last_root = 0;
    if (n_edge > 0) {
        CU_CHECK(ctx, cudaMemcpyAsync(&last_rank, rank.p + (n_edge - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        CU_CHECK(ctx, cudaMemcpyAsync(&last_root, root.p + (n_edge - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    }
    if (out_cam && n_edge > 0) {
        CU_CHECK(ctx, d_img.stage(mem, node_img, (size_t)n_node, s));
        CU_CHECK(ctx, d_xy.stage(mem, node_xy, (size_t)n_node * 2, s));
        CU_CHECK(ctx, o_cam.stage(mem, out_cam, (size_t)n_edge * 2));
        CU_CHECK(ctx, o_lm.stage(mem, out_lm, (size_t)n_edge * 2));
        CU_CHECK(ctx, o_xy.stage(mem, out_xy, (size_t)n_edge * 4));
        k_mg_flatten<<<div_up(n_edge, T), T, 0, s>>>(n_edge, d_a.d, d_b.d, d_img.d, d_xy.d, o_label.d, o_cam.d, o_lm.d, o_xy.d);
        KERNEL_POST(ctx);
        CU_CHECK(ctx, o_cam.finish(s)); CU_CHECK(ctx, o_lm.finish(s)); CU_CHECK(ctx, o_xy.finish(s));
    }
    CU_CHECK(ctx, o_label.finish(s));
    CU_CHECK(ctx, cudaStreamSynchronize(s));
    *out_n_landmark = last_rank + last_root;
    return PTZBA_OK;
}
