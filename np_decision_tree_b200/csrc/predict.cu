// K5: batched tree traversal.
//
// replaces utils.inference / utils.inference_leaf      (ref decision_tree/utils.py:15-31, 34-48)
//
// The reference advances all rows one level per numpy pass: go left iff
// data[row, feature[node]] <= threshold[node]; a row stops where the chosen child is -1.
// Rows are independent, so one thread walks one row to its leaf; the tree (<= 2047 nodes at
// depth 10, 20 bytes per node) is staged in shared memory.  Each visited node costs one
// 32-byte sector of the row (row-major X), so the algorithmic traffic is
// 32 * path_length + 8 (+4 for the leaf id) bytes per row.
#include "common.cuh"

namespace {

struct TreeDev {
    const int *feature;
    const int *left;
    const int *right;
    const double *threshold;
    const double *value;
};

template <bool SMEM_TREE>
__global__ void __launch_bounds__(256) predict_kernel(const double *__restrict__ X, int64_t N, int64_t F, int64_t ldx,
                                                      TreeDev t, int n_nodes, double *__restrict__ out_value,
                                                      int *__restrict__ out_leaf) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int *feature = t.feature, *left = t.left, *right = t.right;
    const double *threshold = t.threshold;
    if (SMEM_TREE) {
        double *s_thr = reinterpret_cast<double *>(smem_raw);
        int *s_feat = reinterpret_cast<int *>(s_thr + n_nodes);
        int *s_left = s_feat + n_nodes;
        int *s_right = s_left + n_nodes;
        for (int i = threadIdx.x; i < n_nodes; i += blockDim.x) {
            s_thr[i] = t.threshold[i];
            s_feat[i] = t.feature[i];
            s_left[i] = t.left[i];
            s_right[i] = t.right[i];
        }
        __syncthreads();
        feature = s_feat;
        left = s_left;
        right = s_right;
        threshold = s_thr;
    }
    for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < N;
         row += (int64_t)gridDim.x * blockDim.x) {
        const double *x = X + row * ldx;
        int node = 0;
        for (int step = 0; step < n_nodes; ++step) {
            const int l = left[node], r = right[node];
            if (l == -1 && r == -1) break;  // utils.py:25-27: no row moves from a leaf
            int f = feature[node];
            if (f < 0) f += (int)F;  // numpy negative index (utils.py:22 reads column F-2 at a leaf)
            const bool go_left = __ldg(x + f) <= threshold[node];  // utils.py:22
            const int nxt = go_left ? l : r;                       // utils.py:23-24
            if (nxt == -1) break;                                  // utils.py:28-30: stay
            DCHECK(nxt >= 0 && nxt < n_nodes && f >= 0 && f < (int)F);
            node = nxt;
        }
        if (out_value) out_value[row] = t.value[node];
        if (out_leaf) out_leaf[row] = node;
    }
}

}  // namespace

int predict_device(b200dt_ctx *ctx, const double *dX, int64_t N, int64_t F, int64_t ldx, const int *d_feature,
                   const int *d_left, const int *d_right, const double *d_threshold, const double *d_value,
                   int n_nodes, double *d_out_value, int *d_out_leaf, double avg_path_hint) {
    if (N <= 0) return 0;
    TreeDev t{d_feature, d_left, d_right, d_threshold, d_value};
    const size_t smem = (size_t)n_nodes * 20 + 16;
    int64_t blocks = (N + 255) / 256;
    const int64_t cap = (int64_t)ctx->sm_count * 64;
    if (blocks > cap) blocks = cap;
    const double bytes = (32.0 * avg_path_hint + (d_out_value ? 8.0 : 0.0) + (d_out_leaf ? 4.0 : 0.0)) * (double)N;
    LaunchScope ls(ctx, KC_PREDICT, bytes);
    if (smem <= 200 * 1024) {
        // (per context = per device: a process may drive several GPUs)
        if (!(ctx->attr_mask & 8u)) {
            CK(cudaFuncSetAttribute(predict_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            ctx->attr_mask |= 8u;
        }
        predict_kernel<true><<<(unsigned)blocks, 256, smem, ctx->stream>>>(dX, N, F, ldx, t, n_nodes, d_out_value, d_out_leaf);
    } else {
        predict_kernel<false><<<(unsigned)blocks, 256, 0, ctx->stream>>>(dX, N, F, ldx, t, n_nodes, d_out_value, d_out_leaf);
    }
    CK(cudaGetLastError());
    return 0;
}
