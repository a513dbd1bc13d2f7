// ba.h - the GPU-resident bundle-adjustment problem (keyframes x ray landmarks x observations) and the
// device-level entry points shared between ba_kernels.cu (per-observation passes), dense.cu (reduced camera
// system factorisation) and ba_solver.cu (trust-region driver).
#pragma once
#include "common.h"
#include "ptz_math.cuh"

// accumulator arena layout (one cudaMemsetAsync clears it): [cost | U N*6 | V M*3 | gc N*3 | gl M*2]
struct BaAccum {
    double* base = nullptr;
    double* cost = nullptr;   // [0] 0.5*sum r^2 is formed on the host from sum r^2 stored here
    double* U = nullptr;      // [N*6] packed (pp,pt,pf,tt,tf,ff), per degree / pixel units
    double* gc = nullptr;     // [N*3]
    double* V = nullptr;      // [M*3] packed (theta-theta, theta-phi, phi-phi)
    double* gl = nullptr;     // [M*2]
    size_t count = 0;
};

struct ptzba_ba {
    ptzba_ctx* ctx = nullptr;
    int n_pose = 0, n_lm = 0;
    int64_t n_obs = 0;
    double u = 0, v = 0;
    bool identity_perm = true;      // caller order already landmark-major
    int max_degree = 0;             // largest number of observations of one landmark
    // observations, landmark-major (SoA, 24 B per observation)
    DevBuf<int32_t> s_cam, s_lm, orig;
    DevBuf<double> s_ox, s_oy;
    DevBuf<int32_t> lm_ptr;         // [M+1] CSR offsets into the sorted arrays
    // second copy in keyframe-major order (landmark ascending inside a keyframe) for the keyframe pass of the fused
    // pass: keyframe blocks accumulate in registers there.
    // the run of every keyframe starts at a multiple of 128 entries (padding: landmark id -1); iter_cam[it] = keyframe of
    // entries [128 it, 128 it + 128)
    DevBuf<int32_t> c_lm, iter_cam;
    DevBuf<double> c_ox, c_oy;
    int n_cam_iter = 0;
    // current parameters
    DevBuf<CamTrig> cam_trig;       // [N]
    DevBuf<LmTrig> lm_trig;         // [M]
    DevBuf<double> accum_store;
    BaAccum acc;
    DevBuf<double> resid;           // [2*n_obs] caller order, staging for host-side residual requests
    DevBuf<double> x_stage, ref_stage;   // persistent H2D staging of x and the reference pose
    // pipelined host-buffer pass (ptzba_ba_normal_equations_begin / ptzba_ba_wait): own copy stream, the events that order it
    // against the compute stream
    cudaStream_t cp_stream = nullptr;
    cudaEvent_t ev_in = nullptr, ev_pass = nullptr, ev_out = nullptr;
    bool async_pending = false;
    double* async_cost_dst = nullptr;
    double* h_cost = nullptr;            // pinned
    ~ptzba_ba() {
        if (cp_stream) cudaStreamDestroy(cp_stream);
        if (ev_in) cudaEventDestroy(ev_in);
        if (ev_pass) cudaEventDestroy(ev_pass);
        if (ev_out) cudaEventDestroy(ev_out);
        if (h_cost) cudaFreeHost(h_cost);
    }
    // ---- solver workspace (allocated on first solve) ----
    DevBuf<double> x_cur, x_trial;          // [3(N-1)+2M] packed parameter vectors
    DevBuf<double> scale_inv;               // [3N + 2M] (camera part incl. slot for pose 0, then landmarks)
    DevBuf<double> Sred;                    // [n x n] reduced camera system, column-major lower, n = 3(N-1)
    DevBuf<double> rhs_c, rhs_l;            // [3N], [2M] right-hand sides / solutions (camera slot 0 unused)
    DevBuf<double> sol_c, sol_l, sol2_c, sol2_l;
    DevBuf<double> Vinv;                    // [M*3] (V + alpha D_l^2)^-1 packed
    DevBuf<double> scal;                    // small device scalar block for reductions
    DevBuf<double> sums_part;               // per-block partial sums of the deterministic k_sums
    DevBuf<unsigned> sums_ticket;
    DevBuf<int> sol_flags;                  // [0] singular V blocks, [1] potrf info
    DevBuf<double> sol_tmp_l, sol_w, sol_dinv;   // back-substitution scratch, D^2 delta, inverted diagonal blocks of chol(S)
    int grid_fused = 0;             // one wave of resident CTAs of the fused pass
    double fused_lm_share = 57.0;   // PTZBA_OPT_FUSED_LM_SHARE: per cent of the per-observation cost that is the landmark-major role's
    // keyframe-pair-major list of observation pairs of this rank's landmark slice, built at the first solve
    int schur_mode = 0;             // PTZBA_SCHUR_* (ptzba_ba_set_option)
    DevBuf<uint32_t> pl_key, pl_val;
    long long pl_n = 0;
    bool pl_ready = false;
    int touch_cam_lo = 0, touch_cam_hi = 0, touch_lm_lo = 0, touch_lm_hi = 0;   // id ranges the observations touch
    // keyframe-sharded mode: compact exchange of the landmarks observed by more than one rank (ptzba_ba_setup_exchange)
    bool exchange_ready = false;
    bool cost_partial = false;      // the cost of the last pass has not been summed over the ranks yet
    int64_t n_shared = 0;
    DevBuf<int32_t> shared_ids;
    DevBuf<double> shared_buf;
    // work partition (ptzba_ba_set_partition); defaults = everything on this rank
    int part_rank = 0, part_world = 1;
    int lm_lo = 0, lm_hi = 0;                  // landmark slice
    int64_t lmo_lo = 0, lmo_hi = 0;            // its observations in the landmark-major list
    int cit_lo = 0, cit_hi = 0;                // slice of the padded keyframe-major list, in iterations of 128 entries
    bool cam_smem = true;           // keyframe trig table fits in shared memory
    bool acc_zeroed = false;        // the arena was cleared by ba_set_params(..., zero_acc = true)
    bool arena_foreign = false;     // a whole-arena all-reduce wrote blocks outside the touched id ranges
};

// ---- device-level passes (all pointers device, enqueued on ba->ctx->stream) ----
int ba_set_params(ptzba_ba* ba, const double* d_x, const double* d_ref_pose3, bool zero_acc = false);   // unpack + trig tables (+ clear ba->acc)
int ba_fused_pass(ptzba_ba* ba, double* d_resid_or_null);                            // -> ba->acc
int ba_residual_pass(ptzba_ba* ba, double* d_resid_or_null, double* d_sumsq, bool reduce = true);   // r (caller order) and sum r^2 (summed over the ranks of a partition unless !reduce)
extern "C" int ptzba_comm_allreduce_f64(ptzba_ctx* ctx, double* device_buf, int64_t count);
