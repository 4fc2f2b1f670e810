// mpcb200.cu -- sm_100a kernels + C-ABI host side of the batched receding-horizon OCP solver (include/mpcb200.h).
//
// Thread mapping.  The stage-parallel phases put one lane on one horizon stage: EVAL and LINESEARCH run one CTA of
// ceil(N/32) warps per instance and read the instance image (the leading, contiguous part of the instance's
// workspace block, mpc_layout.h) from shared memory, where a bulk-async copy staged it; INIT and ASSOCIATE (once per
// solve) run one warp per instance straight on global memory.  Arrays are [component][stage] (stage fastest).
// In the sequential KKT phase ONE LANE owns one instance (mpc_riccati_lane.h): the condensed KKT stage records and
// the Riccati gains live in 32-instance interleaved tiles, a warp sweeps 32 instances in lock-step, and each stage of
// a tile is one contiguous block that a bulk-async copy streams through a shared-memory ring.
//
// This file is the ONLY implementation of the hot path: there is no CPU fallback.  Every entry point fails with
// MPCB200_E_NODEVICE / MPCB200_E_CUDA when no CUDA device is usable.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "mpc_core.h"
#include "mpc_riccati_lane.h"
#include "mpc_stage.h"
#include "mpc_layout.h"

#define FULLMASK 0xffffffffu
#define WARPS_PER_CTA 4
#define REGROUP_EVERY 4    // IPM iterations between two re-assignments of instances to KKT tile slots
#define MAX_IMG_SMEM (227 * 1024 - 2048)  // dynamic shared memory the eval / line-search kernels may request
#define MAX_GROUP_WARPS 4   // warps of the CTA that owns one instance in the eval / line-search kernels (lane per stage)

// ---- warp reductions (fp64 via two 32-bit shuffles each) ------------------------------------------------
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULLMASK, v, o));
    return v;
}
__device__ __forceinline__ double warp_min(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(FULLMASK, v, o));
    return v;
}

struct InputPtrs
{
    const double* x0; const double* xf; const double* u_prev;      // [B][3],[B][3],[B][2]
    const int* obst_count; const int* obst_type; const double* obst_params; int obst_max;
    const int* vp_count; const double* vp_poses; int vp_max;
    const double* x_init;                                          // [B][N][3] or null
    const unsigned char* reinit;                                   // [B] or null
};

// ---- scatter the compact input arrays into an instance block (inputs of instance `src` into the block W) -------
__device__ __forceinline__ void scatter_one(const WsLayout& L, double* W, const InputPtrs& in, int64_t src, int lane)
{
    const int N = L.N;
    if (lane < 3) { AIN(IN_X0 + lane) = in.x0[src * 3 + lane]; AIN(IN_XF + lane) = in.xf[src * 3 + lane]; }
    if (lane < 2) AIN(IN_UPREV + lane) = in.u_prev ? in.u_prev[src * 2 + lane] : 0.0;
    int nob = 0, nvp = 0;
    if (in.obst_count) nob = min(in.obst_count[src], min(in.obst_max, L.M));
    if (in.vp_count) nvp = min(in.vp_count[src], min(in.vp_max, L.V));
    if (lane == 0)
    {
        AIN(IN_NOBST) = (double)nob; AIN(IN_NVP) = (double)nvp;
        AIN(IN_HASXINIT) = in.x_init ? 1.0 : 0.0;
        AIN(IN_REINIT) = (in.reinit && in.reinit[src]) ? 1.0 : 0.0;
    }
    for (int i = lane; i < nob * MPCB200_OBST_STRIDE; i += 32)
        W[L.oOBST + i] = in.obst_params[src * in.obst_max * MPCB200_OBST_STRIDE + i];
    for (int i = lane; i < nob; i += 32) W[L.oOTYPE + i] = (double)in.obst_type[src * in.obst_max + i];
    for (int i = lane; i < nvp * 3; i += 32) W[L.oVP + i] = in.vp_poses[src * in.vp_max * 3 + i];
    if (in.x_init)
        for (int i = lane; i < 3 * N; i += 32) W[L.oXINIT + i] = in.x_init[src * 3 * N + i];
}

__global__ void scatter_inputs_kernel(WsLayout L, double* ws, int B, InputPtrs in)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B) return;
    scatter_one(L, ws + (int64_t)warp * L.stride, in, warp, lane);
}

// ---- kernel: PHASE_INIT -- cold initial guess or warm-start shift ------------------------------------------
__global__ void init_kernel(Cfg c, WsLayout L, double* ws, int B, int force_cold, int only_new)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B) return;
    double* W = ws + (int64_t)warp * L.stride;
    const int N = L.N;
    if (only_new && ASC(MPCB200_SC_NEW) == 0.0) return;  // streaming: only the slots that were just refilled
    const bool cold = force_cold || ASC(MPCB200_SC_COLD) != 0.0 || AIN(IN_REINIT) != 0.0;
    __syncwarp();
    if (cold)
    {
        for (int k = lane; k < N; k += 32) init_cold_stage(c, L, W, k);
        __syncwarp();
        double nx, ny;
        if (bump_enabled(c, L, W) && bump_normal(L, W, &nx, &ny))
        {
            // choice of the cold initial guess: candidates one after the other, their stages spread over the lanes
            double best = 1e300, best_a = 0.0;
            for (int m = -c.initial_guess_bumps; m <= c.initial_guess_bumps; ++m)
            {
                const double A = BUMP_STEP * (double)m;
                double v = 0.0;
                for (int k = lane; k < N; k += 32) v += bump_stage_violation(c, L, W, k, A, nx, ny);
                const double score = 1e-3 * fabs(A) + warp_sum(v);
                if (bump_better(score, best)) { best = score; best_a = A; }
            }
            for (int k = lane; k < N; k += 32)
                if (k >= 1 && k <= N - 2) { const double o = bump_offset(N, k, best_a); AX(0, k) += o * nx; AX(1, k) += o * ny; }
            __syncwarp();
            if (bump_align_headings(c, best_a))
                for (int k = lane; k < N; k += 32)
                    if (k >= 1 && k <= N - 2) AX(2, k) = bump_heading(L, W, k);
        }
        if (lane == 0) { ASC(MPCB200_SC_DT) = c.dt_ref; ASC(MPCB200_SC_COLD) = 2.0; /* 2: cold init done, repair pending */ }
    }
    else
    {
        if (lane == 0)
        {
            if (c.warm_start && !c.variable_dt) warm_shift_serial(c, L, W);
            else
            {
                for (int i = 0; i < 3; ++i) AX(i, 0) = AIN(IN_X0 + i);
                for (int i = 0; i < 3; ++i)
                    if (c.xf_fixed[i]) AX(i, N - 1) = AIN(IN_XF + i);
            }
            ASC(MPCB200_SC_COLD) = 0.0;
        }
    }
}

// ---- kernel: PHASE_ASSOCIATE -- obstacle / via-point association, initial-guess repair, dual initialisation ----
__global__ void associate_kernel(Cfg c, WsLayout L, double* ws, int B, double uprev_dt, int first_outer, int only_new)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B) return;
    double* W = ws + (int64_t)warp * L.stride;
    const int N = L.N;
    if (only_new && ASC(MPCB200_SC_NEW) == 0.0) return;  // streaming: only the slots that were just refilled
    const bool repair = first_outer && ASC(MPCB200_SC_COLD) == 2.0;
    __syncwarp();
    for (int k = lane; k < N; k += 32) associate_stage(c, L, W, k);
    // via-points: MinTimeViaPointsCost::update with findClosestPose (argmin over the grid, first minimum wins)
    if (has_viapoints(c))
    {
        const int nvp = (int)AIN(IN_NVP);
        int start_idx = 0;
        for (int j = 0; j < nvp && j < L.V; ++j)
        {
            const double vx = W[L.oVP + 3 * j], vy = W[L.oVP + 3 * j + 1];
            double best = 1e300; int bidx = -1;
            for (int i = start_idx + lane; i < N - 1; i += 32)
            {
                const double dx = AX(0, i) - vx, dy = AX(1, i) - vy;
                const double d = sqrt(dx * dx + dy * dy);
                if (d < best) { best = d; bidx = i; }
            }
            // warp argmin with smallest index on ties
            for (int o = 16; o > 0; o >>= 1)
            {
                const double ob = __shfl_xor_sync(FULLMASK, best, o);
                const int oi = __shfl_xor_sync(FULLMASK, bidx, o);
                if (ob < best || (ob == best && oi >= 0 && (bidx < 0 || oi < bidx))) { best = ob; bidx = oi; }
            }
            {
                const double dx = AX(0, N - 1) - vx, dy = AX(1, N - 1) - vy;
                const double d = sqrt(dx * dx + dy * dy);
                if (d < best) { best = d; bidx = N - 1; }
            }
            int idx = bidx;
            if (c.vp_ordered) start_idx = idx + 2;
            if (idx > N - 2) idx = N - 2;
            if (idx < 1) idx = c.vp_ordered ? 1 : -1;
            if (lane == 0) W[L.oVPST + j] = (double)idx;
        }
        for (int j = nvp + lane; j < L.V; j += 32) W[L.oVPST + j] = -1.0;
    }
    __syncwarp();
    if (repair)
    {
        for (int k = lane; k < N; k += 32) project_stage(c, L, W, k);
        __syncwarp();
        {
            // step 2 of the repair: stages in order, the lateral candidates of a pinched stage spread over the lanes
            double nx, ny;
            lateral_normal(L, W, &nx, &ny);
            for (int k = 1; k <= N - 2; ++k)
            {
                if (!lateral_needed(L, W, k)) continue;  // warp-uniform
                const double o_prev = lateral_offset(L, W, k - 1, nx, ny);
                double best = 1e300;
                int best_m = 0;
                for (int m = -LAT_MAX_STEPS + lane; m <= LAT_MAX_STEPS; m += 32)
                {
                    const double cost = lateral_candidate(c, L, W, k, m, o_prev, nx, ny);
                    if (cost < best) { best = cost; best_m = m; }
                }
                for (int o = 16; o > 0; o >>= 1)
                {
                    const double oc = __shfl_xor_sync(FULLMASK, best, o);
                    const int om = __shfl_xor_sync(FULLMASK, best_m, o);
                    if (oc < best || (oc == best && om < best_m)) { best = oc; best_m = om; }
                }
                if (lane == 0) lateral_apply(L, W, k, best_m, best < 1e299, nx, ny);
                __syncwarp();
            }
        }
        for (int k = lane; k < N; k += 32) init_controls_stage(c, L, W, k);
        __syncwarp();
        if (lane == 0) clip_rates_serial(c, L, W, uprev_dt);
        __syncwarp();
    }
    double mu = c.mu_init;
    if (!(mu > 0.0))
    {
        double obj = 0.0, rows = 0.0;
        for (int k = lane; k < N; k += 32) auto_mu_stage(c, L, W, uprev_dt, k, &obj, &rows);
        mu = auto_mu(warp_sum(obj), warp_sum(rows));
    }
    for (int k = lane; k < N; k += 32) init_duals_stage(c, L, W, uprev_dt, k, mu);
    __syncwarp();
    if (lane == 0)
    {
        ASC(MPCB200_SC_MU) = mu; ASC(MPCB200_SC_RHO) = 1.0; ASC(MPCB200_SC_DELTA) = 0.0; ASC(MPCB200_SC_DELTA_LAST) = 0.0;
        ASC(MPCB200_SC_ITER) = 0.0; ASC(MPCB200_SC_NREG) = 0.0; ASC(MPCB200_SC_NBT) = 0.0;
        ASC(MPCB200_SC_DDT) = 0.0; ASC(MPCB200_SC_ALPHA) = 0.0; ASC(MPCB200_SC_TINY) = 0.0; ASC(MPCB200_SC_DEFER) = 0.0;
        if (repair) ASC(MPCB200_SC_COLD) = 0.0;
        if (only_new)
        {
            // streaming: this kernel runs beside the iteration kernels of the other slots; the slot stays parked (status >= 0,
            // never -1 in between) until the next refill kernel, which runs after this one on the main stream, activates it
            ASC(MPCB200_SC_STATUS) = (double)MPCB200_STATUS_INVALID_INPUT;
            ASC(MPCB200_SC_NEW) = 2.0;
        }
        else { ASC(MPCB200_SC_STATUS) = -1.0; ASC(MPCB200_SC_NEW) = 0.0; }
    }
}

// ---- kernel: regroup -- assign tile slots so that instances in the same state share tiles ------------------------------
// key 0: active, last factorisation needed no regularisation; key 1: active, regularised last time (will sweep more
// than once); key 2: finished.  The lane-per-instance KKT kernel pays max-over-lanes per warp, so homogeneous tiles
// remove most of the divergence (and tiles of finished instances exit immediately).  Results do not depend on the slot.
__global__ void __launch_bounds__(1024) regroup_kernel(WsLayout L, const double* ws, int B, int* slot_of, int* inst_of_slot, int nslots)
{
    __shared__ int wcount[3][32];   // per-warp counts of the current chunk -> exclusive scan
    __shared__ int chunk_total[3];
    __shared__ int running[3];      // next free slot of each key class
    __shared__ int total[3];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid < 3) total[tid] = 0;
    __syncthreads();
    // pass 0: class sizes
    for (int c0 = 0; c0 < B; c0 += 1024)
    {
        const int b = c0 + tid;
        int key = -1;
        if (b < B)
        {
            const double* W = ws + (int64_t)b * L.stride;
            key = ASC(MPCB200_SC_STATUS) >= 0.0 ? 2 : (ASC(MPCB200_SC_DELTA_LAST) > 0.0 ? 1 : 0);
        }
#pragma unroll
        for (int k = 0; k < 3; ++k)
        {
            const unsigned m = __ballot_sync(FULLMASK, key == k);
            if (lane == 0 && m) atomicAdd(&total[k], __popc(m));
        }
    }
    __syncthreads();
    if (tid == 0) { running[0] = 0; running[1] = total[0]; running[2] = total[0] + total[1]; }
    __syncthreads();
    // pass 1: deterministic slot assignment (instance order preserved inside a class)
    for (int c0 = 0; c0 < B; c0 += 1024)
    {
        const int b = c0 + tid;
        int key = -1;
        if (b < B)
        {
            const double* W = ws + (int64_t)b * L.stride;
            key = ASC(MPCB200_SC_STATUS) >= 0.0 ? 2 : (ASC(MPCB200_SC_DELTA_LAST) > 0.0 ? 1 : 0);
        }
        int myoff = 0;
#pragma unroll
        for (int k = 0; k < 3; ++k)
        {
            const unsigned m = __ballot_sync(FULLMASK, key == k);
            if (key == k) myoff = __popc(m & ((1u << lane) - 1u));
            if (lane == 0) wcount[k][wid] = __popc(m);
        }
        __syncthreads();
        if (wid < 3)
        {
            const int v = wcount[wid][lane];
            int incl = v;
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULLMASK, incl, o); if (lane >= o) incl += t; }
            wcount[wid][lane] = incl - v;
            if (lane == 31) chunk_total[wid] = incl;
        }
        __syncthreads();
        if (key >= 0)
        {
            const int sidx = running[key] + wcount[key][wid] + myoff;
            slot_of[b] = sidx;
            inst_of_slot[sidx] = b;
        }
        __syncthreads();
        if (tid < 3) running[tid] += chunk_total[tid];
        __syncthreads();
    }
    for (int sidx = B + tid; sidx < nslots; sidx += 1024) inst_of_slot[sidx] = -1;
}

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}


// ---- instance image: the leading part of an instance's workspace staged in shared memory ---------------------------
// The eval / line-search kernels walk long chains of dependent loads (row slots -> obstacle indices -> obstacle
// parameters ...); out of L2 every link costs ~600 cycles.  One elected thread therefore pulls the image -- scalars,
// inputs, iterate, (step,) obstacles: 20-30 KB, contiguous by construction of the layout (mpc_layout.h) -- into
// shared memory with two or three bulk-async (TMA) copies, and the stage functions read it from there.
// Layout of the dynamic shared memory: [0,8) mbarrier, [16, 16 + 8*img_words) image.
struct ImageCopy { int step_src; };  // -1: no step (eval); else offset of the step to place at L.oSTEP (line search)
__device__ __forceinline__ double* load_image(unsigned char* smem, const WsLayout& L, const double* Gp, int img_words, int step_src, int tid)
{
    double* img = reinterpret_cast<double*>(smem + 16);
    const uint32_t bar = smem_addr(smem), dst = smem_addr(img);
    if (tid == 0)
    {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0)
    {
        // scalars of this instance may have been written by this thread a moment ago (generic proxy): order them first
        __threadfence();
        asm volatile("fence.proxy.async;" ::: "memory");
        const uint32_t bytes_a = (uint32_t)L.oSTEP * 8u, bytes_c = (uint32_t)(img_words - L.oOTYPE) * 8u;
        const uint32_t bytes_b = step_src >= 0 ? (uint32_t)(8 * L.N) * 8u : 0u;
        mbar_expect_tx(bar, bytes_a + bytes_b + bytes_c);
        bulk_g2s(dst, Gp, bytes_a, bar);
        if (step_src >= 0) bulk_g2s(dst + (uint32_t)L.oSTEP * 8u, Gp + step_src, bytes_b, bar);
        bulk_g2s(dst + (uint32_t)L.oOTYPE * 8u, Gp + L.oOTYPE, bytes_c, bar);
    }
    mbar_wait(bar, 0);
    return img;
}

// ---- kernel: PHASE_EVAL -----------------------------------------------------------------------------------
// ONE CTA PER INSTANCE, one lane per horizon stage: ceil(N/32) warps (at most MAX_GROUP_WARPS, then the stage loop
// wraps).  Splitting an instance over several warps halves the dependent instruction stream each warp walks through --
// at BASELINE batch sizes every kernel of the IPM iteration is latency- not throughput-bound.
__device__ __forceinline__ void evalacc_warp_reduce(EvalAcc& a)
{
    a.dual_inf = warp_max(a.dual_inf); a.prim_inf = warp_max(a.prim_inf);
    a.sl_max = warp_max(a.sl_max); a.sl_min = warp_min(a.sl_min);
    a.sum_nu = warp_sum(a.sum_nu); a.sum_lam = warp_sum(a.sum_lam); a.inf1 = warp_sum(a.inf1); a.blog = warp_sum(a.blog);
    a.gt0 = warp_sum(a.gt0); a.gt1 = warp_sum(a.gt1); a.gldt = warp_sum(a.gldt); a.htt = warp_sum(a.htt);
    a.obj = warp_sum(a.obj); a.m_ineq = warp_sum(a.m_ineq); a.m_eq = warp_sum(a.m_eq);
}

template <int NW, bool LINES>
__global__ void __maxnreg__(NW <= 2 ? 144 : 168) eval_kernel(Cfg c, WsLayout L, double* ws, double* kkt_tiles, const int* slot_of, int B, double uprev_dt, int* n_active,
                                                             int img_words)
{
    extern __shared__ __align__(128) unsigned char img_smem[];
    __shared__ EvalAcc s_acc[MAX_GROUP_WARPS];
    __shared__ double s_mu;
    __shared__ int s_fin;
    const int inst = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
    double* Gp = ws + (int64_t)inst * L.stride;
    const int N = L.N;
    if (Gp[L.oSCAL + MPCB200_SC_STATUS] >= 0.0) return;  // finished instance: exact no-op (uniform over the CTA)
    double* W = load_image(img_smem, L, Gp, img_words, -1, tid);
    const int slot = slot_of[inst];
    double* Kb = kkt_tiles + (size_t)(slot >> 5) * N * KW * TILE + (slot & 31);
    EvalAcc a;
    evalacc_init(a);
    for (int k = tid; k < N; k += blockDim.x) eval_stage<LINES>(c, L, W, Gp, Kb, uprev_dt, k, a);
    evalacc_warp_reduce(a);
    if (lane == 0) s_acc[wid] = a;
    __syncthreads();
    if (tid == 0)
    {
        for (int w = 1; w < nw; ++w) evalacc_merge(a, s_acc[w]);
        int fin = 0;
        s_mu = eval_finish(c, L, Gp, a, true, &fin);
        s_fin = fin;
        if (!fin && n_active) atomicAdd(n_active, 1);
    }
    __syncthreads();
    if (s_fin) return;
    const double mu = s_mu;
    for (int k = tid; k < N; k += blockDim.x) eval_finalize_stage(L, W, Kb, k, mu);
}

// ---- kernel: PHASE_KKT -- Riccati factorisation + solve, ONE LANE PER INSTANCE ------------------------------------
// One warp per 32-instance tile.  The stage records (backward sweep) and the gains + dynamics (forward sweep) of a tile
// are contiguous 10.5 KB / 10.5-16 KB blocks, so they are streamed through a KKT_RING-deep shared-memory ring with
// 1-D bulk-async copies (cp.async.bulk, the TMA engine) completing on mbarriers: the lanes read their words from
// shared memory (conflict-free, 8 B per lane) and never stall on a global load.  Lanes whose instance is finished (or
// beyond B) idle through the sweeps.
#define KKT_RING 4
struct StepOut
{
    double* W; int oSTEP; int N;
    __device__ void operator()(int k, int c, double v) const { W[oSTEP + c * N + k] = v; }
};

struct SmemView
{
    const double* p;  // word 0 of this lane in the staged block, words TILE apart
    __device__ double operator()(int f) const { return p[f * TILE]; }
};
template <int NG>
struct FwdViewSmem
{
    const double* p;  // staged block: NG gain words, then record words 20..31, then 39..41
    __device__ double gain(int w) const { return p[w * TILE]; }
    __device__ double rec(int f) const { return p[(f < MPCB200_K_D ? NG + (f - MPCB200_K_A) : NG + 12 + (f - MPCB200_K_D)) * TILE]; }
};

template <bool EXT>
struct TmaFeed
{
    static constexpr int NC = EXT ? 5 : 1;
    static constexpr int NG = 25 + 5 * NC;                                             // gain words read by the forward sweep
    static constexpr int FWD_WORDS = NG + 12 + (EXT ? 3 : 0);
    static constexpr int BUF_WORDS = FWD_WORDS > MPCB200_KKT_WORDS ? FWD_WORDS : MPCB200_KKT_WORDS;  // per ring slot, x TILE doubles
    TileRec rec;              // this lane's words in the tile (direct global access: terminal record)
    TileRic ric;              // this lane's gain words (written by the backward sweep)
    const double* rec_tile;   // tile bases for the bulk copies
    const double* ric_tile;
    double* ring;             // shared-memory ring (generic address)
    uint32_t ring_s, bar_s;   // ... and its shared-space address, the mbarriers
    int lane, dt_free;
    uint32_t issued, consumed;

    __device__ bool any(bool p) const { return __any_sync(FULLMASK, p) != 0; }
    __device__ void issue_bwd(int k)
    {
        if (lane == 0)
        {
            const uint32_t slot = issued % KKT_RING, bar = bar_s + 8 * slot;
            mbar_expect_tx(bar, MPCB200_KKT_WORDS * TILE * 8);
            bulk_g2s(ring_s + slot * (BUF_WORDS * TILE * 8), rec_tile + (size_t)k * MPCB200_KKT_WORDS * TILE, MPCB200_KKT_WORDS * TILE * 8, bar);
        }
        ++issued;
    }
    __device__ void issue_fwd(int k)
    {
        if (lane == 0)
        {
            const uint32_t slot = issued % KKT_RING, bar = bar_s + 8 * slot, dst = ring_s + slot * (BUF_WORDS * TILE * 8);
            const bool with_d = EXT && dt_free;
            mbar_expect_tx(bar, (NG + 12 + (with_d ? 3 : 0)) * TILE * 8);
            bulk_g2s(dst, ric_tile + (size_t)k * RICW_MAX * TILE, NG * TILE * 8, bar);
            bulk_g2s(dst + NG * TILE * 8, rec_tile + ((size_t)k * MPCB200_KKT_WORDS + MPCB200_K_A) * TILE, 12 * TILE * 8, bar);
            if (with_d) bulk_g2s(dst + (NG + 12) * TILE * 8, rec_tile + ((size_t)k * MPCB200_KKT_WORDS + MPCB200_K_D) * TILE, 3 * TILE * 8, bar);
        }
        ++issued;
    }
    __device__ const double* acquire()
    {
        const uint32_t slot = consumed % KKT_RING;
        mbar_wait(bar_s + 8 * slot, (consumed / KKT_RING) & 1u);
        return ring + (size_t)slot * BUF_WORDS * TILE + lane;
    }
    __device__ void bwd_start(int kfirst)
    {
        for (int j = 0; j < KKT_RING && kfirst - j >= 0; ++j) issue_bwd(kfirst - j);
    }
    __device__ SmemView bwd_acquire(int) { return SmemView{acquire()}; }
    __device__ void bwd_release(int k)
    {
        __syncwarp();  // every lane is done with the slot before the next copy lands in it
        ++consumed;
        if (k - KKT_RING >= 0) issue_bwd(k - KKT_RING);
    }
    __device__ void bwd_abort()
    {
        while (consumed != issued) { acquire(); ++consumed; }  // drain the copies in flight
    }
    __device__ void fwd_start(int N)
    {
        // the gains were written with ordinary stores by the lanes of this warp: order them before the async-proxy reads
        __threadfence();
        asm volatile("fence.proxy.async;" ::: "memory");
        __syncwarp();
        for (int j = 0; j < KKT_RING && j <= N - 2; ++j) issue_fwd(j);
        n_stages = N - 1;
    }
    int n_stages;
    __device__ FwdViewSmem<NG> fwd_acquire(int) { return FwdViewSmem<NG>{acquire()}; }
    __device__ void fwd_release(int k)
    {
        __syncwarp();
        ++consumed;
        if (k + KKT_RING < n_stages) issue_fwd(k + KKT_RING);
    }
};

template <bool EXT>
__global__ void __launch_bounds__(32) kkt_lane_kernel(Cfg c, WsLayout L, double* ws, const double* kkt_tiles, double* ric_tiles, size_t ric_attempt_stride,
                                                       const int* inst_of_slot, int B, int spec, unsigned long long* counters)
{
    extern __shared__ __align__(128) unsigned char kkt_smem[];
    const int lane = threadIdx.x & 31;
    const int tile = blockIdx.x;
    const int attempt = blockIdx.y;  // speculative mode: attempt 0 and 1 of the regularisation schedule run side by side
    const int N = L.N;
    const int b = inst_of_slot[tile * TILE + lane];
    double* W = b >= 0 ? ws + (int64_t)b * L.stride : ws;
    const bool active = b >= 0 && !(ASC(MPCB200_SC_STATUS) >= 0.0);
    if (!__any_sync(FULLMASK, active)) return;
    uint64_t* bars = reinterpret_cast<uint64_t*>(kkt_smem);
    TmaFeed<EXT> feed;
    feed.rec_tile = kkt_tiles + (size_t)tile * N * KW * TILE;
    feed.ric_tile = ric_tiles + (size_t)attempt * ric_attempt_stride + (size_t)tile * N * RICW_MAX * TILE;
    feed.rec = TileRec{feed.rec_tile + lane};
    feed.ric = TileRic{const_cast<double*>(feed.ric_tile) + lane};
    feed.ring = reinterpret_cast<double*>(kkt_smem + 128);
    feed.ring_s = smem_addr(feed.ring);
    feed.bar_s = smem_addr(bars);
    feed.lane = lane; feed.dt_free = c.variable_dt;
    feed.issued = feed.consumed = 0; feed.n_stages = N - 1;
    if (lane == 0)
    {
        for (int i = 0; i < KKT_RING; ++i) mbar_init(feed.bar_s + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    StepOut step{W, attempt ? L.oSTEP2 : L.oSTEP, N};
    double ddt = 0.0, delta = 0.0;
    int nreg = 0;
    const double htt = active ? ASC(MPCB200_SC_HTT) : 0.0, gt = active ? ASC(MPCB200_SC_GT) : 0.0, dlast = active ? ASC(MPCB200_SC_DELTA_LAST) : 0.0;
    const int ok = riccati_solve_lane<EXT>(c, N, feed, step, active, htt, gt, dlast, spec ? attempt : 0, spec ? 1 : MAX_INERTIA_TRIES, &ddt, &delta, &nreg);
    if (!active) return;
    if (counters)
    {
        if (attempt == 0) atomicAdd(counters, 1ull);
        atomicAdd(counters + 1, spec ? 1ull : (unsigned long long)(nreg + (ok ? 1 : 0)));
    }
    if (spec)
    {
        // the line-search kernel picks the winner (kkt_resolve)
        if (attempt == 0) { ASC(MPCB200_SC_KKT_OK0) = (double)ok; ASC(MPCB200_SC_DELTA) = delta; ASC(MPCB200_SC_DDT) = ddt; }
        else { ASC(MPCB200_SC_KKT_OK1) = (double)ok; ASC(MPCB200_SC_DELTA1) = delta; ASC(MPCB200_SC_DDT1) = ddt; }
        return;
    }
    ASC(MPCB200_SC_NREG) += (double)nreg;
    if (!ok && delta <= MAX_DELTA)
    {
        // factorisation budget of this iteration spent: null step, the next iteration resumes at this delta (DELTA_LAST / 3)
        ASC(MPCB200_SC_DELTA_LAST) = 3.0 * delta;
        ASC(MPCB200_SC_DEFER) = 1.0;
        return;
    }
    if (!ok) { ASC(MPCB200_SC_STATUS) = (double)MPCB200_STATUS_NUMERICAL_ERROR; return; }
    ASC(MPCB200_SC_DEFER) = 0.0;
    ASC(MPCB200_SC_DDT) = ddt;
    ASC(MPCB200_SC_DELTA) = delta;
    ASC(MPCB200_SC_DELTA_LAST) = delta;
}

template <bool EXT>
static size_t kkt_smem_bytes() { return 128 + (size_t)KKT_RING * TmaFeed<EXT>::BUF_WORDS * TILE * 8; }

// ---- kernel: PHASE_LINESEARCH -------------------------------------------------------------------------------
// one CTA per instance (lane per stage, see eval_kernel); thread 0 owns the scalar decisions of the line search.
struct LsShared
{
    LsAcc acc[MAX_GROUP_WARPS];
    TrialAcc tr[MAX_GROUP_WARPS];
    int hist[CLIP_BINS + 1];
    double alpha, a_dual;
    int accept;
};

#define GSC(i_) Gp[L.oSCAL + (i_)]
template <bool LINES>
__global__ void __launch_bounds__(MAX_GROUP_WARPS * 32, 4) linesearch_kernel(Cfg c, WsLayout L, double* ws, const double* kkt_tiles, const int* slot_of, int B, double uprev_dt,
                                                                             int spec, int img_words)
{
    extern __shared__ __align__(128) unsigned char img_smem[];
    __shared__ LsShared sh;
    __shared__ int s_win;
    const int inst = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
    double* Gp = ws + (int64_t)inst * L.stride;
    const int N = L.N;
    if (GSC(MPCB200_SC_STATUS) >= 0.0) return;
    int step_src = L.oSTEP;
    if (spec)
    {
        // the KKT phase ran attempts 0 and 1 of the regularisation schedule side by side: pick the winner
        __syncthreads();
        if (tid == 0)
        {
            int nreg = 0;
            double dnext = 0.0;
            const double d0 = GSC(MPCB200_SC_DELTA), d1 = GSC(MPCB200_SC_DELTA1);
            const int win = kkt_resolve(GSC(MPCB200_SC_KKT_OK0) != 0.0, GSC(MPCB200_SC_KKT_OK1) != 0.0, d0, d1, GSC(MPCB200_SC_DELTA_LAST), &nreg, &dnext);
            GSC(MPCB200_SC_NREG) += (double)nreg;
            if (win == -2) GSC(MPCB200_SC_STATUS) = (double)MPCB200_STATUS_NUMERICAL_ERROR;
            else if (win == -1) { GSC(MPCB200_SC_DELTA_LAST) = 3.0 * dnext; GSC(MPCB200_SC_DEFER) = 1.0; }
            else
            {
                const double dw = win ? d1 : d0;
                GSC(MPCB200_SC_DEFER) = 0.0;
                if (win) GSC(MPCB200_SC_DDT) = GSC(MPCB200_SC_DDT1);
                GSC(MPCB200_SC_DELTA) = dw;
                GSC(MPCB200_SC_DELTA_LAST) = dw;
            }
            s_win = win;
        }
        __syncthreads();
        if (s_win == -2) return;
        if (s_win == 1) step_src = L.oSTEP2;
    }
    if (GSC(MPCB200_SC_DEFER) != 0.0)
    {
        // the KKT phase spent its factorisation budget: null step
        __syncthreads();
        if (tid == 0) { GSC(MPCB200_SC_DEFER) = 0.0; GSC(MPCB200_SC_ITER) += 1.0; GSC(MPCB200_SC_ALPHA) = 0.0; }
        return;
    }
    double* W = load_image(img_smem, L, Gp, img_words, step_src, tid);  // image incl. the winning step at L.oSTEP
    const int slot = slot_of[inst];
    const double* Kb = kkt_tiles + (size_t)(slot >> 5) * N * KW * TILE + (slot & 31);
    LsAcc a;
    lsacc_init(a);
    // histogram of the blocking step ratios (+ row count) -> threshold bin of the clipped rows -> primal step length
    for (int j = tid; j <= CLIP_BINS; j += blockDim.x) sh.hist[j] = 0;
    __syncthreads();
    for (int k = tid; k < N; k += blockDim.x) ls_stage_steps(c, L, W, Gp, Kb, uprev_dt, k, a, sh.hist);
    __syncthreads();
    const int jt = clip_threshold_bin(sh.hist, sh.hist[CLIP_BINS]);  // same value in every thread
    for (int k = tid; k < N; k += blockDim.x) a.a_p = fmin(a.a_p, ls_stage_ap(L, W, k, jt));
    a.a_p = warp_min(a.a_p); a.a_d = warp_min(a.a_d);
    a.dphi_bar = warp_sum(a.dphi_bar); a.curv = warp_sum(a.curv); a.dJ = warp_sum(a.dJ);
    if (lane == 0) sh.acc[wid] = a;
    __syncthreads();
    // scalars of the merit function (thread 0 only)
    double mu = 0.0, rho = 1.0, phi0 = 0.0, dphi = 0.0, a_d = 1.0;
    if (tid == 0)
    {
        for (int w = 1; w < nw; ++w)
        {
            const LsAcc& o = sh.acc[w];
            a.a_p = fmin(a.a_p, o.a_p); a.a_d = fmin(a.a_d, o.a_d);
            a.dphi_bar += o.dphi_bar; a.curv += o.curv; a.dJ += o.dJ;
        }
        mu = ASC(MPCB200_SC_MU);
        const double inf1 = ASC(MPCB200_SC_INF), obj = ASC(MPCB200_SC_OBJ), blog = ASC(MPCB200_SC_BLOG);
        const double num = a.dJ + a.dphi_bar + 0.5 * (a.curv > 0 ? a.curv : 0.0);
        if (inf1 > 1e-14)
        {
            const double rho_trial = num / ((1.0 - 0.1) * inf1);
            if (rho < rho_trial) rho = rho_trial + 1.0;
        }
        phi0 = obj - mu * blog + rho * inf1;
        dphi = a.dJ + a.dphi_bar - rho * inf1;
        a_d = a.a_d;
        sh.alpha = a.a_p;
    }
    __syncthreads();
    double alpha = sh.alpha;
    int nbt = 0;
    for (int bt = 0; bt < MAX_BACKTRACK; ++bt)
    {
        TrialAcc t;
        t.obj = t.inf1 = t.blog = 0.0;
        for (int k = tid; k < N; k += blockDim.x) ls_stage_trial<LINES>(c, L, W, Gp, uprev_dt, k, alpha, t);
        t.obj = warp_sum(t.obj); t.inf1 = warp_sum(t.inf1); t.blog = warp_sum(t.blog);
        if (lane == 0) sh.tr[wid] = t;
        __syncthreads();
        if (tid == 0)
        {
            for (int w = 1; w < nw; ++w) { t.obj += sh.tr[w].obj; t.inf1 += sh.tr[w].inf1; t.blog += sh.tr[w].blog; }
            const double phi = t.obj - mu * t.blog + rho * t.inf1;
            sh.accept = (phi <= phi0 + ARMIJO * alpha * dphi || (bt > 0 && fabs(phi - phi0) <= 1e-13 * (1.0 + fabs(phi0)))) ? 1 : 0;
        }
        __syncthreads();
        const int accept = sh.accept;
        __syncthreads();  // sh.accept / sh.tr are rewritten by the next trial
        if (accept) break;
        alpha *= 0.5;
        ++nbt;
    }
    if (tid == 0) sh.a_dual = a_d > alpha ? alpha : a_d;
    __syncthreads();
    const double a_dual = sh.a_dual;
    for (int k = tid; k < N; k += blockDim.x) ls_stage_update<LINES>(c, L, W, Gp, uprev_dt, k, alpha, a_dual);
    if (tid == 0)
    {
        if (c.variable_dt) GSC(MPCB200_SC_DT) = ASC(MPCB200_SC_DT) + alpha * ASC(MPCB200_SC_DDT);
        GSC(MPCB200_SC_ALPHA) = alpha;
        GSC(MPCB200_SC_RHO) = rho;
        GSC(MPCB200_SC_ITER) = ASC(MPCB200_SC_ITER) + 1.0;
        GSC(MPCB200_SC_NBT) = ASC(MPCB200_SC_NBT) + (double)nbt;
        const double tiny = alpha < TINY_STEP ? ASC(MPCB200_SC_TINY) + 1.0 : 0.0;
        GSC(MPCB200_SC_TINY) = tiny;
        if (tiny >= (double)TINY_STEP_COUNT) GSC(MPCB200_SC_STATUS) = (double)MPCB200_STATUS_NUMERICAL_ERROR;  /* jammed: give up */
    }
}

// ---- kernel: gather results into compact arrays ----------------------------------------------------------------
struct OutputPtrs { double* u_seq; double* x_seq; double* dt; int* status; double* kkt; int* iters; double* u_packed; };
__device__ __forceinline__ void gather_one(const WsLayout& L, const double* W, const OutputPtrs& o, int64_t dst, int lane)
{
    const int N = L.N;
    for (int k = lane; k < N; k += 32)
    {
        const int kk = k <= N - 2 ? k : N - 2;
        o.u_seq[(dst * N + k) * 2 + 0] = AU(0, kk);
        o.u_seq[(dst * N + k) * 2 + 1] = AU(1, kk);
        o.x_seq[(dst * N + k) * 3 + 0] = AX(0, k);
        o.x_seq[(dst * N + k) * 3 + 1] = AX(1, k);
        o.x_seq[(dst * N + k) * 3 + 2] = normalize_theta(AX(2, k));
        if (k <= N - 2)
        {
            o.u_packed[(dst * (N - 1) + k) * 2 + 0] = AU(0, k);
            o.u_packed[(dst * (N - 1) + k) * 2 + 1] = AU(1, k);
        }
    }
    if (lane == 0)
    {
        o.dt[dst] = ASC(MPCB200_SC_DT);
        const double st = ASC(MPCB200_SC_STATUS);
        o.status[dst] = st < 0 ? MPCB200_STATUS_MAX_ITER : (int)st;
        o.kkt[dst] = ASC(MPCB200_SC_ERR0);
        o.iters[dst] = (int)ASC(MPCB200_SC_ITER);
    }
}

__global__ void gather_outputs_kernel(WsLayout L, const double* ws, int B, OutputPtrs o)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B) return;
    gather_one(L, ws + (int64_t)warp * L.stride, o, warp, lane);
}

// ---- streaming: a pool of B slots works through a queue of instances (continuous batching) ---------------------------
// A slot whose instance has finished hands its result over (gather into the arrays of the whole job at the instance's
// index) and takes the next instance of the queue; the IPM iteration kernels neither know nor care which slot holds
// which instance.  Run every STREAM_REFILL_EVERY iterations, followed by the init / associate kernels restricted to the
// slots marked SC_NEW.  counters[0] = next instance of the queue, counters[1] = results handed over.
#define STREAM_REFILL_EVERY 2
struct StreamState { int* slot_inst; int* counters; int total; };
__global__ void stream_begin_kernel(WsLayout L, double* ws, int B, StreamState st)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b == 0) { st.counters[0] = 0; st.counters[1] = 0; }
    if (b >= B) return;
    double* W = ws + (int64_t)b * L.stride;
    ASC(MPCB200_SC_STATUS) = (double)MPCB200_STATUS_INVALID_INPUT;  // any value >= 0: the slot is free
    ASC(MPCB200_SC_NEW) = 0.0;
    st.slot_inst[b] = -1;
}
__global__ void stream_refill_kernel(WsLayout L, double* ws, int B, InputPtrs in, OutputPtrs o, StreamState st)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B) return;
    double* W = ws + (int64_t)warp * L.stride;
    if (!(ASC(MPCB200_SC_STATUS) >= 0.0)) return;  // still iterating
    const double nw = ASC(MPCB200_SC_NEW);
    __syncwarp();
    if (nw != 0.0)
    {
        // refilled in the previous round, initialised on the side stream since then: starts iterating now
        if (lane == 0 && nw == 2.0) { ASC(MPCB200_SC_STATUS) = -1.0; ASC(MPCB200_SC_NEW) = 0.0; }
        return;
    }
    const int id = st.slot_inst[warp];
    __syncwarp();
    if (id >= 0)
    {
        gather_one(L, W, o, id, lane);
        if (lane == 0) atomicAdd(&st.counters[1], 1);
    }
    int nid = 0;
    if (lane == 0) nid = atomicAdd(&st.counters[0], 1);
    nid = __shfl_sync(FULLMASK, nid, 0);
    if (nid < st.total)
    {
        scatter_one(L, W, in, nid, lane);
        if (lane == 0) { st.slot_inst[warp] = nid; ASC(MPCB200_SC_COLD) = 1.0; ASC(MPCB200_SC_NEW) = 1.0; }
    }
    else if (lane == 0)
    {
        st.slot_inst[warp] = -1;
        if (nid > (1 << 30)) st.counters[0] = st.total;  // idle slots keep asking: never let the counter wrap
    }
}

__global__ void reset_kernel(WsLayout L, double* ws, int B, const unsigned char* which)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    if (which && !which[b]) return;
    double* W = ws + (int64_t)b * L.stride;
    ASC(MPCB200_SC_COLD) = 1.0;
    ASC(MPCB200_SC_STATUS) = -1.0;
}

// ---- horizon change (grid adaptation): pack the warm trajectories, switch the layout, resample into the new one ----
// record per instance: SCAL words, X (3 x n_old), U (2 x n_old)
__global__ void resample_pack_kernel(WsLayout L, const double* ws, double* rec, int rec_words, int B)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double* W = ws + (int64_t)b * L.stride;
    double* r = rec + (size_t)b * rec_words;
    const int N = L.N;
    for (int i = 0; i < MPCB200_SCAL_WORDS; ++i) r[i] = W[L.oSCAL + i];
    for (int i = 0; i < 3 * N; ++i) r[MPCB200_SCAL_WORDS + i] = W[L.oX + i];
    for (int i = 0; i < 2 * N; ++i) r[MPCB200_SCAL_WORDS + 3 * N + i] = W[L.oU + i];
}
__global__ void resample_unpack_kernel(WsLayout L, double* ws, const double* rec, int rec_words, int n_old, int B)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double* W = ws + (int64_t)b * L.stride;
    const double* r = rec + (size_t)b * rec_words;
    for (int i = 0; i < MPCB200_SCAL_WORDS; ++i) W[L.oSCAL + i] = r[i];
    if (ASC(MPCB200_SC_COLD) != 0.0) return;  // empty grid: the next step initialises it at the new horizon
    ASC(MPCB200_SC_DT) = resample_serial(n_old, r + MPCB200_SCAL_WORDS, r + MPCB200_SCAL_WORDS + 3 * n_old, r[MPCB200_SC_DT], L.N, W + L.oX, W + L.oU);
}

// ---- costmap -> point obstacles (MpcLocalPlannerROS::updateObstacleContainerWithCostmap, mpc_local_planner_ros.cpp:474-499) ----
// The reference walks the cells mx = 0..size_x-2 (outer), my = 0..size_y-2 (inner), keeps the LETHAL ones that are not farther
// than behind_dist behind the robot and appends them as point obstacles at the cell centres.  HBM-bound byte work, one byte per
// cell, read ONCE:
//   mark    a thread owns four adjacent columns (one 32-bit load per row, a warp reads 128 contiguous bytes), tests the word
//           for a LETHAL byte with one bit trick, applies the filter to the few hits and records them as one bit per cell in
//           per-column masks (32 rows per word, 1/8 byte per cell) next to the per-column counts;
//   offsets exclusive scan of the column counts of each robot;
//   emit    a thread owns one column and walks its mask words in row order: column offsets + bit order reproduce the
//           reference's push_back order (mx outer, my inner) exactly.
#define COSTMAP_LETHAL 254u   // costmap_2d::LETHAL_OBSTACLE
struct CostmapArgs
{
    int size_x, size_y;
    double resolution, behind_dist;
    const unsigned char* cost;   // [B][size_y][size_x]
    const double* origin;        // [B][2]
    const double* pose;          // [B][3]
};
// Costmap2D::mapToWorld: cell centre
__device__ __forceinline__ double costmap_world(double o, int m, double res) { return o + ((double)m + 0.5) * res; }
__device__ __forceinline__ bool costmap_keep(const CostmapArgs& a, int mx, int my, double ox, double oy, double px, double py, double dirx, double diry)
{
    const double dx = costmap_world(ox, mx, a.resolution) - px, dy = costmap_world(oy, my, a.resolution) - py;
    // "not far behind the robot" (mpc_local_planner_ros.cpp:492-493)
    return !(dx * dirx + dy * diry < 0.0 && sqrt(dx * dx + dy * dy) > a.behind_dist);
}
__device__ __forceinline__ bool word_has_lethal(unsigned w)
{
    const unsigned x = w ^ 0xFEFEFEFEu;                       // LETHAL bytes become zero bytes
    return ((x - 0x01010101u) & ~x & 0x80808080u) != 0u;
}
template <bool VEC>   // VEC: size_x % 4 == 0, every row of every map starts 4-byte aligned
__global__ void costmap_mark_kernel(CostmapArgs a, int B, int nrb, int Wp, unsigned* mask /*[B][nrb][Wp]*/, int* colcount /*[B][size_x]*/)
{
    const int b = blockIdx.y;
    const int c0 = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
    if (b >= B || c0 >= a.size_x) return;
    const unsigned char* map = a.cost + (size_t)b * a.size_x * a.size_y;
    const double ox = a.origin[2 * b], oy = a.origin[2 * b + 1];
    const double px = a.pose[3 * b], py = a.pose[3 * b + 1];
    double diry, dirx;
    sincos(a.pose[3 * b + 2], &diry, &dirx);   // PoseSE2::orientationUnitVec
    int cnt[4] = {0, 0, 0, 0};
    const int rows = a.size_y - 1;
    for (int rb = 0; rb < nrb; ++rb)
    {
        unsigned m[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int r0 = 0; r0 < 32; r0 += 8)
        {
            unsigned w[8];
#pragma unroll
            for (int r = 0; r < 8; ++r)
            {
                const int my = rb * 32 + r0 + r;
                w[r] = 0u;
                if (my < rows)
                {
                    const unsigned char* q = map + (size_t)my * a.size_x + c0;
                    if (VEC) w[r] = *reinterpret_cast<const unsigned*>(q);
                    else
                        for (int i = 0; i < 4; ++i)
                            if (c0 + i < a.size_x) w[r] |= (unsigned)q[i] << (8 * i);
                }
            }
#pragma unroll
            for (int r = 0; r < 8; ++r)
            {
                if (!word_has_lethal(w[r])) continue;
                const int my = rb * 32 + r0 + r;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (((w[r] >> (8 * i)) & 0xFFu) == COSTMAP_LETHAL && c0 + i < a.size_x - 1 &&
                        costmap_keep(a, c0 + i, my, ox, oy, px, py, dirx, diry))
                        m[i] |= 1u << (r0 + r);
            }
        }
        *reinterpret_cast<uint4*>(mask + ((size_t)b * nrb + rb) * Wp + c0) = make_uint4(m[0], m[1], m[2], m[3]);
#pragma unroll
        for (int i = 0; i < 4; ++i) cnt[i] += __popc(m[i]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (c0 + i < a.size_x) colcount[(size_t)b * a.size_x + c0 + i] = cnt[i];
}
__global__ void costmap_emit_kernel(CostmapArgs a, int B, int nrb, int Wp, const unsigned* mask, const int* colstart, int max_out,
                                    double* params /*[B][max_out][MPCB200_OBST_STRIDE]*/, int* type /*[B][max_out]*/)
{
    const int b = blockIdx.y;
    const int mx = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B || mx >= a.size_x - 1) return;
    const double ox = a.origin[2 * b], oy = a.origin[2 * b + 1];
    int o = colstart[(size_t)b * a.size_x + mx];
    for (int rb = 0; rb < nrb && o < max_out; ++rb)
    {
        unsigned m = mask[((size_t)b * nrb + rb) * Wp + mx];
        while (m && o < max_out)
        {
            const int my = rb * 32 + __ffs(m) - 1;
            m &= m - 1u;
            double* q = params + ((size_t)b * max_out + o) * MPCB200_OBST_STRIDE;
            q[0] = costmap_world(ox, mx, a.resolution); q[1] = costmap_world(oy, my, a.resolution);
            for (int i = 2; i < MPCB200_OBST_STRIDE; ++i) q[i] = 0.0;
            type[(size_t)b * max_out + o] = MPCB200_OBST_POINT;
            ++o;
        }
    }
}
// exclusive scan of the column counts of one robot (one CTA per robot); found = total, count = min(total, max_out)
__global__ void costmap_offsets_kernel(int size_x, int B, const int* colcount, int* colstart, int max_out, int* count, int* found)
{
    const int b = blockIdx.x;
    __shared__ int carry;
    __shared__ int warp_tot[32];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int base = 0; base < size_x; base += blockDim.x)
    {
        const int i = base + threadIdx.x;
        const int v = i < size_x ? colcount[(size_t)b * size_x + i] : 0;
        int incl = v;
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULLMASK, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) warp_tot[wid] = incl;
        __syncthreads();
        int woff = 0, tot = 0;
        for (int w = 0; w < nw; ++w) { if (w < wid) woff += warp_tot[w]; tot += warp_tot[w]; }
        if (i < size_x) colstart[(size_t)b * size_x + i] = carry + woff + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) { found[b] = carry; count[b] = carry < max_out ? carry : max_out; }
}

__global__ void flush_kernel(double* buf, size_t n)
{
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    for (; i < n; i += (size_t)gridDim.x * blockDim.x) buf[i] = buf[i] * 0.999 + 1.0;
}

// =================================================================================================================
// host side
// =================================================================================================================
struct mpcb200_handle
{
    Cfg cfg;
    WsLayout L;
    int max_batch, device, B;
    int n_cap;                  // horizon the buffers were sized for at create (mpcb200_resample moves cfg.n within [3, n_cap])
    double* d_resample;         // scratch of mpcb200_resample, allocated on first use
    void* d_cm; size_t cm_cap; double costmap_ms;  // scratch of mpcb200_costmap_obstacles (grown on demand), device ms of its last call
    double* ws;
    double *kkt_tiles, *ric_tiles;
    size_t ric_attempt_stride;  // doubles between the gain tiles of KKT attempt 0 and 1 (speculative mode)
    int num_sms;
    int spec_mode;              // MPCB200_OPT_KKT_ATTEMPTS: 0 auto, 1 serial, 2 side by side
    int refill_every;           // MPCB200_OPT_STREAM_REFILL_EVERY: IPM iterations between two refills of the streaming pool
    int spec;                   // this solve runs the two KKT attempts of an iteration side by side (small batches)
    unsigned timing_mask;       // phases bracketed by CUDA events inside solve (bit = phase id); default: KKT only
    cudaStream_t stream, own_stream;  // stream in use / the stream the handle created
    // compact device input / output staging
    double *d_x0, *d_xf, *d_uprev, *d_obst, *d_vp, *d_xinit;
    int *d_obst_count, *d_obst_type, *d_vp_count;
    unsigned char* d_reinit;
    double *d_useq, *d_xseq, *d_dt, *d_kkt, *d_upacked;
    int *d_status, *d_iters, *d_nactive, *d_slot_of, *d_inst_of_slot;
    unsigned long long* d_counters;
    int* h_nactive;  // pinned, two poll slots
    int* nactive_ptr;            // where the next eval launch counts unfinished instances (or null)
    cudaEvent_t poll_ev[2];
    double* d_flush; size_t flush_n;
    int has_obst, has_vp, has_xinit, has_reinit, obst_max, vp_max;
    int only_new;   // streaming: init / associate touch only the slots marked SC_NEW
    // streaming job: inputs / outputs of the whole queue on the device (grown on demand), slot -> instance map, counters
    size_t stream_cap;
    double *s_x0, *s_xf, *s_uprev, *s_obst, *s_vp, *s_useq, *s_xseq, *s_dt, *s_kkt, *s_upacked;
    int *s_obst_count, *s_obst_type, *s_vp_count, *s_status, *s_iters, *d_slot_inst, *d_stream_counters;
    int* h_stream_counters;  // pinned, two poll slots
    cudaStream_t side_stream;            // streaming: init / associate of refilled slots run beside the iterations of the others
    cudaEvent_t ev_refill, ev_ready;
    int has_lines;  // line obstacles in the batch, moving obstacles or midpoint differences: the eval / line-search kernels are launched with those (rarely used) paths compiled in
    double uprev_dt;
    mpcb200_stats stats;
    std::vector<cudaEvent_t> ev;  // pool of event pairs
    std::vector<int> ev_phase;
    size_t ev_used;
    std::string err;
};

static std::string g_create_err = "";

static int set_err(mpcb200_handle* h, int code, const std::string& msg)
{
    if (h) h->err = msg; else g_create_err = msg;
    return code;
}
#define CK(call)                                                                                                   \
    do {                                                                                                           \
        cudaError_t e_ = (call);                                                                                   \
        if (e_ != cudaSuccess)                                                                                     \
            return set_err(h, MPCB200_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));                 \
    } while (0)

extern "C" void mpcb200_default_config(mpcb200_config* c)
{
    memset(c, 0, sizeof(*c));
    c->robot_type = MPCB200_ROBOT_UNICYCLE;
    c->wheelbase = 0.5; c->length_rear = 1.0; c->length_front = 1.0;
    c->u_lb[0] = -0.2; c->u_lb[1] = -0.3; c->u_ub[0] = 0.4; c->u_ub[1] = 0.3;
    c->du_lb[0] = c->du_lb[1] = -MPCB200_INF; c->du_ub[0] = c->du_ub[1] = MPCB200_INF;
    c->n = 20; c->dt_ref = 0.3; c->variable_dt = 1; c->dt_lb = 0.0; c->dt_ub = 10.0;
    c->xf_fixed[0] = c->xf_fixed[1] = c->xf_fixed[2] = 1;
    c->collocation = MPCB200_COLLOC_FORWARD; c->warm_start = 1;
    c->objective = MPCB200_OBJ_MINIMUM_TIME;
    c->vp_position_weight = 1.0; c->vp_orientation_weight = 0.0;
    c->min_obstacle_dist = 0.5; c->force_inclusion_dist = 0.5; c->cutoff_dist = 2.0;
    c->footprint_type = MPCB200_FOOTPRINT_POINT;
    c->k_max_obstacles_per_stage = 5;
    c->max_iter = 100; c->tol = 1e-6; c->mu_init = 0.0; c->outer_iterations = 1; c->quadratic_integral_form = 0;
    c->initial_guess_bumps = 4;
    c->enable_dynamic_obstacles = 0;
    c->terminal_ball = 0; c->terminal_ball_gamma = 5.0;
    for (int i = 0; i < 9; ++i) c->terminal_ball_S[i] = (i % 4 == 0) ? 1.0 : 0.0;
    c->cost_integration = MPCB200_COST_LEFT_SUM;
    c->hybrid_cost_minimum_time = 0;
}

static int validate_config(const mpcb200_config* c, std::string& why)
{
    if (c->n < 3 || c->n > 512) { why = "n must be in [3, 512]"; return MPCB200_E_INVALID; }
    if (c->robot_type < 0 || c->robot_type > 3) { why = "unknown robot_type"; return MPCB200_E_INVALID; }
    if (c->collocation != MPCB200_COLLOC_FORWARD && c->collocation != MPCB200_COLLOC_MIDPOINT)
    { why = "collocation: forward_differences and midpoint_differences are implemented, crank_nicolson is not"; return MPCB200_E_UNSUPPORTED; }
    if (c->objective < 0 || c->objective > 2) { why = "unknown objective"; return MPCB200_E_INVALID; }
    if (c->cost_integration != MPCB200_COST_LEFT_SUM && c->cost_integration != MPCB200_COST_TRAPEZOIDAL)
    { why = "unknown cost_integration"; return MPCB200_E_INVALID; }
    if (c->footprint_type < 0 || c->footprint_type > 4) { why = "unknown footprint_type"; return MPCB200_E_INVALID; }
    if (c->footprint_type == MPCB200_FOOTPRINT_POLYGON && (c->n_poly < 1 || c->n_poly > MPCB200_MAX_POLY))
    { why = "polygon footprint needs 1..16 vertices"; return MPCB200_E_INVALID; }
    if (c->k_max_obstacles_per_stage < 0 || c->k_max_obstacles_per_stage > 16) { why = "k_max_obstacles_per_stage must be in [0,16]"; return MPCB200_E_INVALID; }
    if (!(c->dt_ref > 0)) { why = "dt_ref must be > 0"; return MPCB200_E_INVALID; }
    if (c->variable_dt && !(c->dt_ub > c->dt_lb)) { why = "dt_ub must exceed dt_lb"; return MPCB200_E_INVALID; }
    if (has_mintime(*c) && !c->variable_dt) { why = "minimum_time objectives need variable_dt"; return MPCB200_E_INVALID; }
    if (!(c->tol > 0) || c->max_iter < 1) { why = "tol > 0 and max_iter >= 1 required"; return MPCB200_E_INVALID; }
    for (int i = 0; i < 2; ++i)
        if (!(c->u_ub[i] > c->u_lb[i])) { why = "u_ub must exceed u_lb"; return MPCB200_E_INVALID; }
    return 0;
}


extern "C" int mpcb200_create(const mpcb200_config* cfg, int max_batch, int device, mpcb200_handle** out)
{
    mpcb200_handle* h = nullptr;
    if (!cfg || !out || max_batch < 1) return set_err(nullptr, MPCB200_E_INVALID, "bad arguments");
    std::string why;
    int rc = validate_config(cfg, why);
    if (rc) return set_err(nullptr, rc, why);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return set_err(nullptr, MPCB200_E_NODEVICE, std::string("no CUDA device (") + cudaGetErrorString(e) + "): this solver has no CPU fallback");
    if (device < 0 || device >= ndev) return set_err(nullptr, MPCB200_E_INVALID, "device index out of range");
    h = new mpcb200_handle();
    h->cfg = *cfg; h->max_batch = max_batch; h->device = device; h->B = 0; h->ws = nullptr; h->kkt_tiles = nullptr; h->ric_tiles = nullptr; h->ev_used = 0;
    memset(&h->stats, 0, sizeof(h->stats));
    make_layout(cfg, MAX_OBST, MAX_VP, h->L);
    h->n_cap = cfg->n; h->d_resample = nullptr; h->d_cm = nullptr; h->cm_cap = 0; h->costmap_ms = 0.0;
    h->uprev_dt = 0.0; h->has_obst = h->has_vp = h->has_xinit = h->has_reinit = 0; h->obst_max = h->vp_max = 0;
#define CKC(call)                                                                                                  \
    do {                                                                                                           \
        cudaError_t e_ = (call);                                                                                   \
        if (e_ != cudaSuccess) { std::string m = std::string(#call) + ": " + cudaGetErrorString(e_); delete h; return set_err(nullptr, MPCB200_E_CUDA, m); } \
    } while (0)
    CKC(cudaSetDevice(device));
    CKC(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
    h->stream = h->own_stream;
    const size_t B = (size_t)max_batch, N = (size_t)cfg->n;
    CKC(cudaMalloc(&h->ws, B * h->L.stride * sizeof(double)));
    CKC(cudaMemsetAsync(h->ws, 0, B * h->L.stride * sizeof(double), h->stream));
    {
        const size_t ntiles = (B + TILE - 1) / TILE;
        CKC(cudaMalloc(&h->kkt_tiles, ntiles * N * KW * TILE * sizeof(double)));
        h->ric_attempt_stride = ntiles * N * RICW_MAX * TILE;
        CKC(cudaMalloc(&h->ric_tiles, 2 * h->ric_attempt_stride * sizeof(double)));
        CKC(cudaMemsetAsync(h->kkt_tiles, 0, ntiles * N * KW * TILE * sizeof(double), h->stream));
        CKC(cudaMemsetAsync(h->ric_tiles, 0, 2 * h->ric_attempt_stride * sizeof(double), h->stream));
        h->spec = 0; h->spec_mode = 0; h->refill_every = STREAM_REFILL_EVERY; h->timing_mask = 1u << MPCB200_PHASE_KKT;
        CKC(cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device));
    }
    CKC(cudaMalloc(&h->d_x0, B * 3 * 8)); CKC(cudaMalloc(&h->d_xf, B * 3 * 8)); CKC(cudaMalloc(&h->d_uprev, B * 2 * 8));
    CKC(cudaMalloc(&h->d_obst, B * MAX_OBST * MPCB200_OBST_STRIDE * 8)); CKC(cudaMalloc(&h->d_obst_count, B * 4));
    CKC(cudaMalloc(&h->d_obst_type, B * MAX_OBST * 4));
    CKC(cudaMalloc(&h->d_vp, B * MAX_VP * 3 * 8)); CKC(cudaMalloc(&h->d_vp_count, B * 4));
    CKC(cudaMalloc(&h->d_xinit, B * N * 3 * 8)); CKC(cudaMalloc(&h->d_reinit, B));
    CKC(cudaMalloc(&h->d_useq, B * N * 2 * 8)); CKC(cudaMalloc(&h->d_xseq, B * N * 3 * 8)); CKC(cudaMalloc(&h->d_dt, B * 8));
    CKC(cudaMalloc(&h->d_kkt, B * 8)); CKC(cudaMalloc(&h->d_upacked, B * (N - 1) * 2 * 8));
    CKC(cudaMalloc(&h->d_status, B * 4)); CKC(cudaMalloc(&h->d_iters, B * 4)); CKC(cudaMalloc(&h->d_nactive, 8));
    CKC(cudaMalloc(&h->d_slot_of, B * 4)); CKC(cudaMalloc(&h->d_inst_of_slot, ((B + TILE - 1) / TILE) * TILE * 4));
    CKC(cudaMalloc(&h->d_counters, 16)); CKC(cudaMemsetAsync(h->d_counters, 0, 16, h->stream));
    CKC(cudaFuncSetAttribute(eval_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_IMG_SMEM));
    CKC(cudaFuncSetAttribute(eval_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_IMG_SMEM));
    CKC(cudaFuncSetAttribute(eval_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_IMG_SMEM));
    CKC(cudaFuncSetAttribute(eval_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_IMG_SMEM));
    CKC(cudaFuncSetAttribute(eval_kernel<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_IMG_SMEM));
    CKC(cudaFuncSetAttribute(eval_kernel<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_IMG_SMEM));
    CKC(cudaFuncSetAttribute(eval_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_IMG_SMEM));
    CKC(cudaFuncSetAttribute(eval_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_IMG_SMEM));
    CKC(cudaFuncSetAttribute(linesearch_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_IMG_SMEM));
    CKC(cudaFuncSetAttribute(linesearch_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_IMG_SMEM));
    CKC(cudaFuncSetAttribute(kkt_lane_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kkt_smem_bytes<true>()));
    CKC(cudaFuncSetAttribute(kkt_lane_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kkt_smem_bytes<false>()));
    CKC(cudaMallocHost(&h->h_nactive, 8));
    h->nactive_ptr = nullptr;
    h->only_new = 0; h->stream_cap = 0;
    h->s_x0 = h->s_xf = h->s_uprev = h->s_obst = h->s_vp = h->s_useq = h->s_xseq = h->s_dt = h->s_kkt = h->s_upacked = nullptr;
    h->s_obst_count = h->s_obst_type = h->s_vp_count = h->s_status = h->s_iters = nullptr;
    CKC(cudaMalloc(&h->d_slot_inst, B * 4)); CKC(cudaMalloc(&h->d_stream_counters, 8));
    CKC(cudaMallocHost(&h->h_stream_counters, 16));
    CKC(cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking));
    CKC(cudaEventCreateWithFlags(&h->ev_refill, cudaEventDisableTiming)); CKC(cudaEventCreateWithFlags(&h->ev_ready, cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&h->poll_ev[0], cudaEventDisableTiming)); CKC(cudaEventCreateWithFlags(&h->poll_ev[1], cudaEventDisableTiming));
    h->flush_n = (size_t)40 * 1024 * 1024;  // 320 MB > 126 MB L2
    CKC(cudaMalloc(&h->d_flush, h->flush_n * 8));
    CKC(cudaMemsetAsync(h->d_flush, 0, h->flush_n * 8, h->stream));
    {
        // all instances start cold
        reset_kernel<<<(max_batch + 127) / 128, 128, 0, h->stream>>>(h->L, h->ws, max_batch, nullptr);
        CKC(cudaGetLastError());
    }
    CKC(cudaStreamSynchronize(h->stream));
    *out = h;
    return MPCB200_OK;
}

extern "C" void mpcb200_destroy(mpcb200_handle* h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    void* ptrs[] = {h->ws, h->kkt_tiles, h->ric_tiles, h->d_x0, h->d_xf, h->d_uprev, h->d_obst, h->d_obst_count, h->d_obst_type, h->d_vp, h->d_vp_count, h->d_xinit,
                    h->d_reinit, h->d_useq, h->d_xseq, h->d_dt, h->d_kkt, h->d_upacked, h->d_status, h->d_iters, h->d_nactive, h->d_flush, h->d_counters, h->d_slot_of, h->d_inst_of_slot};
    for (void* p : ptrs) if (p) cudaFree(p);
    void* sptrs[] = {h->s_x0, h->s_xf, h->s_uprev, h->s_obst, h->s_vp, h->s_useq, h->s_xseq, h->s_dt, h->s_kkt, h->s_upacked, h->s_obst_count,
                     h->s_obst_type, h->s_vp_count, h->s_status, h->s_iters, h->d_slot_inst, h->d_stream_counters, h->d_resample, h->d_cm};
    for (void* p : sptrs) if (p) cudaFree(p);
    if (h->h_stream_counters) cudaFreeHost(h->h_stream_counters);
    if (h->h_nactive) cudaFreeHost(h->h_nactive);
    for (auto& e : h->ev) cudaEventDestroy(e);
    cudaStreamDestroy(h->side_stream); cudaEventDestroy(h->ev_refill); cudaEventDestroy(h->ev_ready);
    cudaStreamDestroy(h->own_stream);
    delete h;
}

extern "C" const char* mpcb200_last_error(const mpcb200_handle* h) { return h ? h->err.c_str() : g_create_err.c_str(); }

static inline int grid_for(int B, int wpc) { return (B + wpc - 1) / wpc; }

// event-pair pool: device time per phase launch (CUDA events on the solver stream)
static int ev_begin(mpcb200_handle* h, int phase)
{
    if (h->ev_used + 2 > h->ev.size())
    {
        for (int i = 0; i < 256; ++i) { cudaEvent_t e; if (cudaEventCreate(&e) != cudaSuccess) return -1; h->ev.push_back(e); }
        h->ev_phase.resize(h->ev.size() / 2);
    }
    h->ev_phase[h->ev_used / 2] = phase;
    cudaEventRecord(h->ev[h->ev_used], h->stream);
    return 0;
}
static void ev_end(mpcb200_handle* h)
{
    cudaEventRecord(h->ev[h->ev_used + 1], h->stream);
    h->ev_used += 2;
}
static void ev_collect(mpcb200_handle* h)
{
    for (size_t i = 0; i + 1 < h->ev_used; i += 2)
    {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, h->ev[i], h->ev[i + 1]) == cudaSuccess)
        {
            const int p = h->ev_phase[i / 2];
            h->stats.ms[p] += ms; h->stats.launches[p] += 1;
        }
    }
    h->ev_used = 0;
}

static int launch_phase(mpcb200_handle* h, int phase, int B, int force_cold, int first_outer, bool timed)
{
    const int grid4 = grid_for(B, WARPS_PER_CTA);
    const int gw = (h->cfg.n + 31) / 32;
    const int group_threads = 32 * (gw < MAX_GROUP_WARPS ? gw : MAX_GROUP_WARPS);
    const int spec = h->spec;
    // instance image staged in shared memory by the eval / line-search kernels: everything up to the obstacles in use
    const int m_used = h->has_obst ? ((h->obst_max + 1) & ~1) : 0;
    const int img_words = h->L.oOBST + MPCB200_OBST_STRIDE * m_used;
    const size_t img_smem = 16 + (size_t)img_words * 8;
    if (img_smem > MAX_IMG_SMEM) return set_err(h, MPCB200_E_UNSUPPORTED, "horizon too long: the instance image does not fit in shared memory");
    if (timed && ev_begin(h, phase)) return set_err(h, MPCB200_E_CUDA, "cudaEventCreate failed");
    switch (phase)
    {
        case MPCB200_PHASE_INIT: init_kernel<<<grid4, WARPS_PER_CTA * 32, 0, h->stream>>>(h->cfg, h->L, h->ws, B, force_cold, h->only_new); break;
        case MPCB200_PHASE_ASSOCIATE: associate_kernel<<<grid4, WARPS_PER_CTA * 32, 0, h->stream>>>(h->cfg, h->L, h->ws, B, h->uprev_dt, first_outer, h->only_new); break;
        case MPCB200_PHASE_EVAL:
#define EVAL_LAUNCH(NW, LN) eval_kernel<NW, LN><<<B, NW * 32, img_smem, h->stream>>>(h->cfg, h->L, h->ws, h->kkt_tiles, h->d_slot_of, B, h->uprev_dt, h->nactive_ptr, img_words)
            switch (group_threads >> 5)
            {
                case 1: if (h->has_lines) EVAL_LAUNCH(1, true); else EVAL_LAUNCH(1, false); break;
                case 2: if (h->has_lines) EVAL_LAUNCH(2, true); else EVAL_LAUNCH(2, false); break;
                case 3: if (h->has_lines) EVAL_LAUNCH(3, true); else EVAL_LAUNCH(3, false); break;
                default: if (h->has_lines) EVAL_LAUNCH(4, true); else EVAL_LAUNCH(4, false); break;
            }
#undef EVAL_LAUNCH
            break;
        case MPCB200_PHASE_KKT:
        {
            const bool ext = h->cfg.variable_dt || h->cfg.xf_fixed[0] || h->cfg.xf_fixed[1] || h->cfg.xf_fixed[2];
            const int ntiles = (B + TILE - 1) / TILE;
            const dim3 kgrid(ntiles, spec ? 2 : 1);
            if (ext) kkt_lane_kernel<true><<<kgrid, 32, kkt_smem_bytes<true>(), h->stream>>>(h->cfg, h->L, h->ws, h->kkt_tiles, h->ric_tiles, h->ric_attempt_stride, h->d_inst_of_slot, B, spec, h->d_counters);
            else kkt_lane_kernel<false><<<kgrid, 32, kkt_smem_bytes<false>(), h->stream>>>(h->cfg, h->L, h->ws, h->kkt_tiles, h->ric_tiles, h->ric_attempt_stride, h->d_inst_of_slot, B, spec, h->d_counters);
            break;
        }
        case MPCB200_PHASE_LINESEARCH:
            if (h->has_lines) linesearch_kernel<true><<<B, group_threads, img_smem, h->stream>>>(h->cfg, h->L, h->ws, h->kkt_tiles, h->d_slot_of, B, h->uprev_dt, spec, img_words);
            else linesearch_kernel<false><<<B, group_threads, img_smem, h->stream>>>(h->cfg, h->L, h->ws, h->kkt_tiles, h->d_slot_of, B, h->uprev_dt, spec, img_words);
            break;
        default: return set_err(h, MPCB200_E_INVALID, "unknown phase");
    }
    if (timed) ev_end(h);
    h->stats.launches_total += 1;
    CK(cudaGetLastError());
    return 0;
}

static int launch_phase_eval(mpcb200_handle* h, int B, int* nactive, bool timed)
{
    h->nactive_ptr = nactive;
    const int rc = launch_phase(h, MPCB200_PHASE_EVAL, B, 0, 0, timed);
    h->nactive_ptr = nullptr;
    return rc;
}

static int launch_regroup(mpcb200_handle* h, int B)
{
    const int nslots = ((B + TILE - 1) / TILE) * TILE;
    regroup_kernel<<<1, 1024, 0, h->stream>>>(h->L, h->ws, B, h->d_slot_of, h->d_inst_of_slot, nslots);
    h->stats.launches_total += 1;
    CK(cudaGetLastError());
    return 0;
}

static int check_batch(mpcb200_handle* h, int B)
{
    if (!h) return MPCB200_E_INVALID;
    if (B < 1 || B > h->max_batch) return set_err(h, MPCB200_E_INVALID, "batch size out of range");
    return 0;
}

static int upload_inputs(mpcb200_handle* h, int B, const double* x0, const double* xf, const double* u_prev, double u_prev_dt,
                         const mpcb200_obstacles* obst, const mpcb200_viapoints* vp, const double* x_init, const unsigned char* reinit)
{
    if (!x0 || !xf) return set_err(h, MPCB200_E_INVALID, "x0 and xf are required");
    const size_t N = (size_t)h->cfg.n;
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(h->d_x0, x0, (size_t)B * 3 * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_xf, xf, (size_t)B * 3 * 8, cudaMemcpyHostToDevice, h->stream));
    h->stats.h2d_bytes += (long long)B * 6 * 8;
    if (u_prev) { CK(cudaMemcpyAsync(h->d_uprev, u_prev, (size_t)B * 2 * 8, cudaMemcpyHostToDevice, h->stream)); h->stats.h2d_bytes += (long long)B * 16; }
    else CK(cudaMemsetAsync(h->d_uprev, 0, (size_t)B * 2 * 8, h->stream));
    h->uprev_dt = u_prev_dt;
    h->has_obst = 0; h->obst_max = 0; h->has_lines = is_midpoint(h->cfg);  // the kernel variants with the rarely used paths compiled in
    if (obst && obst->count && obst->max_per_instance > 0)
    {
        if (obst->max_per_instance > MAX_OBST) return set_err(h, MPCB200_E_UNSUPPORTED, "more than 64 obstacles per instance");
        const size_t M = (size_t)obst->max_per_instance;
        int lines = 0;
        for (size_t i = 0; i < (size_t)B * M; ++i)
        {
            if (obst->type[i] < MPCB200_OBST_POINT || obst->type[i] > MPCB200_OBST_LINE) return set_err(h, MPCB200_E_INVALID, "unknown obstacle type");
            lines |= obst->type[i] == MPCB200_OBST_LINE;
        }
        h->has_lines = lines || h->cfg.enable_dynamic_obstacles || is_midpoint(h->cfg);
        CK(cudaMemcpyAsync(h->d_obst_count, obst->count, (size_t)B * 4, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(h->d_obst_type, obst->type, (size_t)B * M * 4, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(h->d_obst, obst->params, (size_t)B * M * MPCB200_OBST_STRIDE * 8, cudaMemcpyHostToDevice, h->stream));
        h->stats.h2d_bytes += (long long)(B * 4 + B * M * 4 + B * M * MPCB200_OBST_STRIDE * 8);
        h->has_obst = 1; h->obst_max = (int)M;
    }
    h->has_vp = 0; h->vp_max = 0;
    if (vp && vp->count && vp->max_per_instance > 0)
    {
        if (vp->max_per_instance > MAX_VP) return set_err(h, MPCB200_E_UNSUPPORTED, "more than 8 via-points per instance");
        const size_t V = (size_t)vp->max_per_instance;
        CK(cudaMemcpyAsync(h->d_vp_count, vp->count, (size_t)B * 4, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(h->d_vp, vp->poses, (size_t)B * V * 3 * 8, cudaMemcpyHostToDevice, h->stream));
        h->stats.h2d_bytes += (long long)(B * 4 + B * V * 24);
        h->has_vp = 1; h->vp_max = (int)V;
    }
    h->has_xinit = 0;
    if (x_init) { CK(cudaMemcpyAsync(h->d_xinit, x_init, (size_t)B * N * 3 * 8, cudaMemcpyHostToDevice, h->stream)); h->has_xinit = 1; h->stats.h2d_bytes += (long long)(B * N * 24); }
    h->has_reinit = 0;
    if (reinit) { CK(cudaMemcpyAsync(h->d_reinit, reinit, (size_t)B, cudaMemcpyHostToDevice, h->stream)); h->has_reinit = 1; h->stats.h2d_bytes += B; }
    InputPtrs in;
    in.x0 = h->d_x0; in.xf = h->d_xf; in.u_prev = h->d_uprev;
    in.obst_count = h->has_obst ? h->d_obst_count : nullptr; in.obst_type = h->d_obst_type; in.obst_params = h->d_obst; in.obst_max = h->obst_max;
    in.vp_count = h->has_vp ? h->d_vp_count : nullptr; in.vp_poses = h->d_vp; in.vp_max = h->vp_max;
    in.x_init = h->has_xinit ? h->d_xinit : nullptr;
    in.reinit = h->has_reinit ? h->d_reinit : nullptr;
    scatter_inputs_kernel<<<grid_for(B, WARPS_PER_CTA), WARPS_PER_CTA * 32, 0, h->stream>>>(h->L, h->ws, B, in);
    h->stats.launches_total += 1;
    CK(cudaGetLastError());
    h->B = B;
    return 0;
}

// the solve proper: INIT, then outer_iterations x (ASSOCIATE, interior-point iterations)
static int solve_device(mpcb200_handle* h, int B, int force_cold, double* solve_time_s)
{
    cudaEvent_t t0, t1;
    CK(cudaEventCreate(&t0)); CK(cudaEventCreate(&t1));
    CK(cudaEventRecord(t0, h->stream));
    const unsigned tm = h->timing_mask;
    auto timed = [&](int phase) { return ((tm >> phase) & 1u) != 0; };
    // small batches leave most SMs idle in the KKT phase: run both regularisation attempts of an iteration side by side
    h->spec = h->spec_mode == 0 ? ((2 * ((B + TILE - 1) / TILE) <= h->num_sms) ? 1 : 0) : (h->spec_mode == 2 ? 1 : 0);
    int rc = launch_phase(h, MPCB200_PHASE_INIT, B, force_cold, 0, timed(MPCB200_PHASE_INIT));
    if (rc) return rc;
    const int outer = h->cfg.outer_iterations > 0 ? h->cfg.outer_iterations : 1;
    for (int oi = 0; oi < outer; ++oi)
    {
        if ((rc = launch_phase(h, MPCB200_PHASE_ASSOCIATE, B, 0, oi == 0, timed(MPCB200_PHASE_ASSOCIATE)))) return rc;
        // The number of unfinished instances is polled every POLL iterations, one poll behind: the host keeps queueing
        // iterations while the count of the previous poll travels back, so the stream never drains (a finished
        // instance makes every kernel an immediate no-op, so the few surplus iterations are free).
        const int POLL = 4;
        int pending = -1;  // slot of the poll in flight
        bool done = false;
        for (int it = 0; it <= h->cfg.max_iter && !done; ++it)
        {
            const bool poll = (it % POLL == POLL - 1) || it == h->cfg.max_iter;
            const int slot = (it / POLL) & 1;
            if (poll) CK(cudaMemsetAsync(h->d_nactive + slot, 0, 4, h->stream));
            if (it % REGROUP_EVERY == 0 && (rc = launch_regroup(h, B))) return rc;  // slots stay valid in between (finished lanes idle)
            if ((rc = launch_phase_eval(h, B, poll ? h->d_nactive + slot : nullptr, timed(MPCB200_PHASE_EVAL)))) return rc;
            if (poll)
            {
                CK(cudaMemcpyAsync(h->h_nactive + slot, h->d_nactive + slot, 4, cudaMemcpyDeviceToHost, h->stream));
                CK(cudaEventRecord(h->poll_ev[slot], h->stream));
                if (pending >= 0)
                {
                    CK(cudaEventSynchronize(h->poll_ev[pending]));
                    if (h->h_nactive[pending] == 0) done = true;
                }
                pending = slot;
            }
            if (it == h->cfg.max_iter || done) break;
            if ((rc = launch_phase(h, MPCB200_PHASE_KKT, B, 0, 0, timed(MPCB200_PHASE_KKT)))) return rc;
            if ((rc = launch_phase(h, MPCB200_PHASE_LINESEARCH, B, 0, 0, timed(MPCB200_PHASE_LINESEARCH)))) return rc;
        }
    }
    OutputPtrs o{h->d_useq, h->d_xseq, h->d_dt, h->d_status, h->d_kkt, h->d_iters, h->d_upacked};
    gather_outputs_kernel<<<grid_for(B, WARPS_PER_CTA), WARPS_PER_CTA * 32, 0, h->stream>>>(h->L, h->ws, B, o);
    h->stats.launches_total += 1;
    CK(cudaGetLastError());
    CK(cudaEventRecord(t1, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, t0, t1));
    if (solve_time_s) *solve_time_s = ms * 1e-3;
    cudaEventDestroy(t0); cudaEventDestroy(t1);
    ev_collect(h);
    return 0;
}

static int fetch_results(mpcb200_handle* h, int B, double* u_seq, double* x_seq, double* dt_out, int* status, double* kkt_err, int* iters)
{
    const size_t N = (size_t)h->cfg.n;
    if (u_seq) { CK(cudaMemcpyAsync(u_seq, h->d_useq, (size_t)B * N * 16, cudaMemcpyDeviceToHost, h->stream)); h->stats.d2h_bytes += (long long)(B * N * 16); }
    if (x_seq) { CK(cudaMemcpyAsync(x_seq, h->d_xseq, (size_t)B * N * 24, cudaMemcpyDeviceToHost, h->stream)); h->stats.d2h_bytes += (long long)(B * N * 24); }
    if (dt_out) { CK(cudaMemcpyAsync(dt_out, h->d_dt, (size_t)B * 8, cudaMemcpyDeviceToHost, h->stream)); h->stats.d2h_bytes += B * 8; }
    if (status) { CK(cudaMemcpyAsync(status, h->d_status, (size_t)B * 4, cudaMemcpyDeviceToHost, h->stream)); h->stats.d2h_bytes += B * 4; }
    if (kkt_err) { CK(cudaMemcpyAsync(kkt_err, h->d_kkt, (size_t)B * 8, cudaMemcpyDeviceToHost, h->stream)); h->stats.d2h_bytes += B * 8; }
    if (iters) { CK(cudaMemcpyAsync(iters, h->d_iters, (size_t)B * 4, cudaMemcpyDeviceToHost, h->stream)); h->stats.d2h_bytes += B * 4; }
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

// ---- streaming solve: `total` instances through the pool of max_batch slots (continuous batching) ----------------------
static int stream_reserve(mpcb200_handle* h, size_t total)
{
    if (total <= h->stream_cap) return 0;
    void* old[] = {h->s_x0, h->s_xf, h->s_uprev, h->s_obst, h->s_vp, h->s_useq, h->s_xseq, h->s_dt, h->s_kkt, h->s_upacked, h->s_obst_count,
                   h->s_obst_type, h->s_vp_count, h->s_status, h->s_iters};
    for (void* p : old) if (p) cudaFree(p);
    const size_t N = (size_t)h->n_cap, T = total;  // sized for the largest horizon the handle can be resampled to
    CK(cudaMalloc(&h->s_x0, T * 3 * 8)); CK(cudaMalloc(&h->s_xf, T * 3 * 8)); CK(cudaMalloc(&h->s_uprev, T * 2 * 8));
    CK(cudaMalloc(&h->s_obst, T * MAX_OBST * MPCB200_OBST_STRIDE * 8)); CK(cudaMalloc(&h->s_obst_count, T * 4)); CK(cudaMalloc(&h->s_obst_type, T * MAX_OBST * 4));
    CK(cudaMalloc(&h->s_vp, T * MAX_VP * 3 * 8)); CK(cudaMalloc(&h->s_vp_count, T * 4));
    CK(cudaMalloc(&h->s_useq, T * N * 2 * 8)); CK(cudaMalloc(&h->s_xseq, T * N * 3 * 8)); CK(cudaMalloc(&h->s_dt, T * 8)); CK(cudaMalloc(&h->s_kkt, T * 8));
    CK(cudaMalloc(&h->s_upacked, T * (N - 1) * 2 * 8)); CK(cudaMalloc(&h->s_status, T * 4)); CK(cudaMalloc(&h->s_iters, T * 4));
    h->stream_cap = total;
    return 0;
}

extern "C" int mpcb200_solve_stream(mpcb200_handle* h, int total, const double* x0, const double* xf, const double* u_prev, double u_prev_dt,
                                    const mpcb200_obstacles* obst, const mpcb200_viapoints* vp, double* u_seq, double* x_seq, double* dt_out,
                                    int* status, double* kkt_err, int* iters, double* solve_time_s)
{
    if (!h) return MPCB200_E_INVALID;
    if (total < 1 || !x0 || !xf) return set_err(h, MPCB200_E_INVALID, "total >= 1, x0 and xf are required");
    if (h->cfg.outer_iterations > 1) return set_err(h, MPCB200_E_UNSUPPORTED, "streaming runs one outer iteration per instance");
    CK(cudaSetDevice(h->device));
    int rc = stream_reserve(h, (size_t)total);
    if (rc) return rc;
    const size_t T = (size_t)total, N = (size_t)h->cfg.n;
    const int B = total < h->max_batch ? total : h->max_batch;  // slots in use
    // ---- inputs of the whole job ----
    CK(cudaMemcpyAsync(h->s_x0, x0, T * 3 * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->s_xf, xf, T * 3 * 8, cudaMemcpyHostToDevice, h->stream));
    h->stats.h2d_bytes += (long long)(T * 6 * 8);
    if (u_prev) { CK(cudaMemcpyAsync(h->s_uprev, u_prev, T * 2 * 8, cudaMemcpyHostToDevice, h->stream)); h->stats.h2d_bytes += (long long)(T * 16); }
    h->uprev_dt = u_prev_dt;
    h->has_obst = 0; h->obst_max = 0; h->has_lines = is_midpoint(h->cfg); h->has_vp = 0; h->vp_max = 0; h->has_xinit = 0; h->has_reinit = 0;
    if (obst && obst->count && obst->max_per_instance > 0)
    {
        if (obst->max_per_instance > MAX_OBST) return set_err(h, MPCB200_E_UNSUPPORTED, "more than 64 obstacles per instance");
        const size_t M = (size_t)obst->max_per_instance;
        int lines = 0;
        for (size_t i = 0; i < T * M; ++i)
        {
            if (obst->type[i] < MPCB200_OBST_POINT || obst->type[i] > MPCB200_OBST_LINE) return set_err(h, MPCB200_E_INVALID, "unknown obstacle type");
            lines |= obst->type[i] == MPCB200_OBST_LINE;
        }
        h->has_lines = lines || h->cfg.enable_dynamic_obstacles || is_midpoint(h->cfg);
        CK(cudaMemcpyAsync(h->s_obst_count, obst->count, T * 4, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(h->s_obst_type, obst->type, T * M * 4, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(h->s_obst, obst->params, T * M * MPCB200_OBST_STRIDE * 8, cudaMemcpyHostToDevice, h->stream));
        h->stats.h2d_bytes += (long long)(T * 4 + T * M * 4 + T * M * MPCB200_OBST_STRIDE * 8);
        h->has_obst = 1; h->obst_max = (int)M;
    }
    if (vp && vp->count && vp->max_per_instance > 0)
    {
        if (vp->max_per_instance > MAX_VP) return set_err(h, MPCB200_E_UNSUPPORTED, "more than 8 via-points per instance");
        const size_t V = (size_t)vp->max_per_instance;
        CK(cudaMemcpyAsync(h->s_vp_count, vp->count, T * 4, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(h->s_vp, vp->poses, T * V * 3 * 8, cudaMemcpyHostToDevice, h->stream));
        h->stats.h2d_bytes += (long long)(T * 4 + T * V * 24);
        h->has_vp = 1; h->vp_max = (int)V;
    }
    InputPtrs in;
    in.x0 = h->s_x0; in.xf = h->s_xf; in.u_prev = u_prev ? h->s_uprev : nullptr;
    in.obst_count = h->has_obst ? h->s_obst_count : nullptr; in.obst_type = h->s_obst_type; in.obst_params = h->s_obst; in.obst_max = h->obst_max;
    in.vp_count = h->has_vp ? h->s_vp_count : nullptr; in.vp_poses = h->s_vp; in.vp_max = h->vp_max;
    in.x_init = nullptr; in.reinit = nullptr;
    OutputPtrs o{h->s_useq, h->s_xseq, h->s_dt, h->s_status, h->s_kkt, h->s_iters, h->s_upacked};
    StreamState st{h->d_slot_inst, h->d_stream_counters, total};
    // ---- the pool ----
    cudaEvent_t t0, t1;
    CK(cudaEventCreate(&t0)); CK(cudaEventCreate(&t1));
    CK(cudaEventRecord(t0, h->stream));
    const unsigned tm = h->timing_mask;
    auto timed = [&](int phase) { return ((tm >> phase) & 1u) != 0; };
    h->spec = h->spec_mode == 0 ? ((2 * ((B + TILE - 1) / TILE) <= h->num_sms) ? 1 : 0) : (h->spec_mode == 2 ? 1 : 0);
    h->B = B;
    stream_begin_kernel<<<(B + 127) / 128, 128, 0, h->stream>>>(h->L, h->ws, B, st);
    h->stats.launches_total += 1;
    const int grid4 = grid_for(B, WARPS_PER_CTA);
    const int refill_every = h->refill_every;
    const long long max_rounds = ((long long)(total + B - 1) / B + 2) * (h->cfg.max_iter + 2 + 3 * refill_every);
    const int POLL = 4;
    int pending = -1;
    bool done = false;
    h->only_new = 1;
    const cudaStream_t main_stream = h->stream;
    for (long long it = 0; it < max_rounds && !done; ++it)
    {
        if (it % refill_every == 0)
        {
            // refill on the main stream (after the side work of the previous round); the cold initialisation and the
            // association of the refilled slots then run on the side stream, beside the next iterations of the other slots
            if (it > 0) CK(cudaStreamWaitEvent(main_stream, h->ev_ready, 0));
            stream_refill_kernel<<<grid4, WARPS_PER_CTA * 32, 0, main_stream>>>(h->L, h->ws, B, in, o, st);
            h->stats.launches_total += 1;
            CK(cudaEventRecord(h->ev_refill, main_stream));
            CK(cudaStreamWaitEvent(h->side_stream, h->ev_refill, 0));
            h->stream = h->side_stream;
            rc = launch_phase(h, MPCB200_PHASE_INIT, B, 0, 0, timed(MPCB200_PHASE_INIT));
            if (!rc) rc = launch_phase(h, MPCB200_PHASE_ASSOCIATE, B, 0, 1, timed(MPCB200_PHASE_ASSOCIATE));
            h->stream = main_stream;
            if (rc) break;
            CK(cudaEventRecord(h->ev_ready, h->side_stream));
            if ((it / refill_every) % POLL == POLL - 1)
            {
                // results handed over so far, polled one poll behind (see solve_device)
                const int slot = (int)((it / refill_every / POLL) & 1);
                CK(cudaMemcpyAsync(h->h_stream_counters + 2 * slot, h->d_stream_counters, 8, cudaMemcpyDeviceToHost, h->stream));
                CK(cudaEventRecord(h->poll_ev[slot], h->stream));
                if (pending >= 0)
                {
                    CK(cudaEventSynchronize(h->poll_ev[pending]));
                    if (h->h_stream_counters[2 * pending + 1] >= total) done = true;
                }
                pending = slot;
            }
        }
        if (done) break;
        if (it % REGROUP_EVERY == 0 && (rc = launch_regroup(h, B))) break;
        if ((rc = launch_phase_eval(h, B, nullptr, timed(MPCB200_PHASE_EVAL)))) break;
        if ((rc = launch_phase(h, MPCB200_PHASE_KKT, B, 0, 0, timed(MPCB200_PHASE_KKT)))) break;
        if ((rc = launch_phase(h, MPCB200_PHASE_LINESEARCH, B, 0, 0, timed(MPCB200_PHASE_LINESEARCH)))) break;
    }
    h->only_new = 0;
    h->stream = main_stream;
    CK(cudaStreamWaitEvent(main_stream, h->ev_ready, 0));
    if (rc) return rc;
    CK(cudaEventRecord(t1, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    int cnt[2];
    CK(cudaMemcpy(cnt, h->d_stream_counters, 8, cudaMemcpyDeviceToHost));
    if (cnt[1] < total) return set_err(h, MPCB200_E_CUDA, "streaming solve ended before every instance was handed over");
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, t0, t1));
    if (solve_time_s) *solve_time_s = ms * 1e-3;
    cudaEventDestroy(t0); cudaEventDestroy(t1);
    ev_collect(h);
    // ---- results of the whole job ----
    if (u_seq) { CK(cudaMemcpyAsync(u_seq, h->s_useq, T * N * 16, cudaMemcpyDeviceToHost, h->stream)); h->stats.d2h_bytes += (long long)(T * N * 16); }
    if (x_seq) { CK(cudaMemcpyAsync(x_seq, h->s_xseq, T * N * 24, cudaMemcpyDeviceToHost, h->stream)); h->stats.d2h_bytes += (long long)(T * N * 24); }
    if (dt_out) { CK(cudaMemcpyAsync(dt_out, h->s_dt, T * 8, cudaMemcpyDeviceToHost, h->stream)); h->stats.d2h_bytes += (long long)(T * 8); }
    if (status) { CK(cudaMemcpyAsync(status, h->s_status, T * 4, cudaMemcpyDeviceToHost, h->stream)); h->stats.d2h_bytes += (long long)(T * 4); }
    if (kkt_err) { CK(cudaMemcpyAsync(kkt_err, h->s_kkt, T * 8, cudaMemcpyDeviceToHost, h->stream)); h->stats.d2h_bytes += (long long)(T * 8); }
    if (iters) { CK(cudaMemcpyAsync(iters, h->s_iters, T * 4, cudaMemcpyDeviceToHost, h->stream)); h->stats.d2h_bytes += (long long)(T * 4); }
    CK(cudaStreamSynchronize(h->stream));
    // the pool's workspaces now hold arbitrary instances of the job: the next step_batch must start cold
    reset_kernel<<<(B + 127) / 128, 128, 0, h->stream>>>(h->L, h->ws, B, nullptr);
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int mpcb200_step_batch(mpcb200_handle* h, int B, const double* x0, const double* xf, const double* u_prev, double u_prev_dt,
                                  const mpcb200_obstacles* obst, const mpcb200_viapoints* vp, const double* x_init,
                                  const unsigned char* reinit, double* u_seq, double* x_seq, double* dt_out, int* status,
                                  double* kkt_err, int* iters, double* solve_time_s)
{
    int rc = check_batch(h, B);
    if (rc) return rc;
    if ((rc = upload_inputs(h, B, x0, xf, u_prev, u_prev_dt, obst, vp, x_init, reinit))) return rc;
    if ((rc = solve_device(h, B, 0, solve_time_s))) return rc;
    return fetch_results(h, B, u_seq, x_seq, dt_out, status, kkt_err, iters);
}

extern "C" int mpcb200_upload_inputs(mpcb200_handle* h, int B, const double* x0, const double* xf, const double* u_prev, double u_prev_dt,
                                     const mpcb200_obstacles* obst, const mpcb200_viapoints* vp, const double* x_init)
{
    int rc = check_batch(h, B);
    if (rc) return rc;
    if ((rc = upload_inputs(h, B, x0, xf, u_prev, u_prev_dt, obst, vp, x_init, nullptr))) return rc;
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int mpcb200_solve_resident(mpcb200_handle* h, int cold, double* solve_time_s)
{
    if (!h || h->B < 1) return set_err(h, MPCB200_E_INVALID, "no resident inputs: call mpcb200_upload_inputs first");
    CK(cudaSetDevice(h->device));
    return solve_device(h, h->B, cold ? 1 : 0, solve_time_s);
}

extern "C" int mpcb200_fetch_results(mpcb200_handle* h, double* u_seq, double* x_seq, double* dt_out, int* status, double* kkt_err, int* iters)
{
    if (!h || h->B < 1) return set_err(h, MPCB200_E_INVALID, "nothing to fetch");
    CK(cudaSetDevice(h->device));
    return fetch_results(h, h->B, u_seq, x_seq, dt_out, status, kkt_err, iters);
}

extern "C" int mpcb200_device_controls(mpcb200_handle* h, void** dev_ptr, long long* n_doubles)
{
    if (!h || h->B < 1) return set_err(h, MPCB200_E_INVALID, "no batch solved yet");
    if (dev_ptr) *dev_ptr = h->d_upacked;
    if (n_doubles) *n_doubles = (long long)h->B * (h->cfg.n - 1) * 2;
    return 0;
}

extern "C" int mpcb200_reset(mpcb200_handle* h, const unsigned char* which, int B)
{
    if (!h) return MPCB200_E_INVALID;
    CK(cudaSetDevice(h->device));
    const int n = which ? B : h->max_batch;
    if (n < 1 || n > h->max_batch) return set_err(h, MPCB200_E_INVALID, "batch size out of range");
    if (which) CK(cudaMemcpyAsync(h->d_reinit, which, (size_t)n, cudaMemcpyHostToDevice, h->stream));
    reset_kernel<<<(n + 127) / 128, 128, 0, h->stream>>>(h->L, h->ws, n, which ? h->d_reinit : nullptr);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int mpcb200_resample(mpcb200_handle* h, int n_new)
{
    if (!h) return MPCB200_E_INVALID;
    if (n_new < 3 || n_new > h->n_cap) return set_err(h, MPCB200_E_INVALID, "resample: horizon must be in [3, n the handle was created with]");
    if (n_new == h->cfg.n) return MPCB200_OK;
    CK(cudaSetDevice(h->device));
    const int B = h->max_batch, n_old = h->cfg.n;
    const int rec_words = MPCB200_SCAL_WORDS + 5 * h->n_cap;
    if (!h->d_resample) CK(cudaMalloc(&h->d_resample, (size_t)B * rec_words * sizeof(double)));
    resample_pack_kernel<<<(B + 127) / 128, 128, 0, h->stream>>>(h->L, h->ws, h->d_resample, rec_words, B);
    CK(cudaGetLastError());
    h->cfg.n = n_new;
    make_layout(&h->cfg, h->L.M, h->L.V, h->L);
    resample_unpack_kernel<<<(B + 127) / 128, 128, 0, h->stream>>>(h->L, h->ws, h->d_resample, rec_words, n_old, B);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
    return MPCB200_OK;
}

extern "C" int mpcb200_get_horizon(const mpcb200_handle* h, int* n, int* n_capacity)
{
    if (!h) return MPCB200_E_INVALID;
    if (n) *n = h->cfg.n;
    if (n_capacity) *n_capacity = h->n_cap;
    return MPCB200_OK;
}

// ---- kernel-level access ------------------------------------------------------------------------------------
static int field_info(const mpcb200_handle* h, int field, int* off, int* cnt)
{
    const WsLayout& L = h->L;
    switch (field)
    {
        case MPCB200_F_X: *off = L.oX; *cnt = 3; return 0;
        case MPCB200_F_U: *off = L.oU; *cnt = 2; return 0;
        case MPCB200_F_NU: *off = L.oNU; *cnt = 3; return 0;
        case MPCB200_F_S: *off = L.oS; *cnt = L.RS; return 0;
        case MPCB200_F_LAM: *off = L.oLAM; *cnt = L.RS; return 0;
        case MPCB200_F_KKT: *off = 0; *cnt = KW; return 0;
        case MPCB200_F_STEP: *off = L.oSTEP; *cnt = 8; return 0;
        case MPCB200_F_SCAL: *off = L.oSCAL; *cnt = MPCB200_SCAL_WORDS; return 0;
        case MPCB200_F_OBSIDX: *off = L.oOBS; *cnt = L.K > 0 ? L.K : 1; return 0;
    }
    return -1;
}

extern "C" int mpcb200_ws_count(const mpcb200_handle* h, int field)
{
    int off, cnt;
    if (!h || field_info(h, field, &off, &cnt)) return MPCB200_E_INVALID;
    return cnt;
}

extern "C" int mpcb200_ws_read(mpcb200_handle* h, int field, int B, double* dst)
{
    int off, cnt, rc = check_batch(h, B);
    if (rc) return rc;
    if (!dst || field_info(h, field, &off, &cnt)) return set_err(h, MPCB200_E_INVALID, "bad field");
    CK(cudaSetDevice(h->device));
    const int N = h->L.N;
    if (field == MPCB200_F_KKT)
    {   // device layout: 32-instance interleaved tiles [tile][k][42][32]; the API presents [B][42][N]
        const size_t ntiles = ((size_t)B + TILE - 1) / TILE, tw = (size_t)N * KW * TILE;
        std::vector<double> tmp(ntiles * tw);
        std::vector<int> slot(B);
        CK(cudaMemcpyAsync(tmp.data(), h->kkt_tiles, tmp.size() * 8, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(slot.data(), h->d_slot_of, (size_t)B * 4, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        for (int b = 0; b < B; ++b)
        {
            const int sl = slot[b];
            for (int k = 0; k < N; ++k)
                for (int f = 0; f < KW; ++f)
                    dst[((size_t)b * KW + f) * N + k] = tmp[(size_t)(sl / TILE) * tw + ((size_t)k * KW + f) * TILE + (sl % TILE)];
        }
        return 0;
    }
    const size_t words = (field == MPCB200_F_SCAL) ? (size_t)cnt : (size_t)cnt * N;
    CK(cudaMemcpy2DAsync(dst, words * 8, h->ws + off, (size_t)h->L.stride * 8, words * 8, (size_t)B, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int mpcb200_ws_write(mpcb200_handle* h, int field, int B, const double* src)
{
    int off, cnt, rc = check_batch(h, B);
    if (rc) return rc;
    if (!src || field_info(h, field, &off, &cnt)) return set_err(h, MPCB200_E_INVALID, "bad field");
    CK(cudaSetDevice(h->device));
    const int N = h->L.N;
    if (field == MPCB200_F_KKT)
    {
        const size_t ntiles = ((size_t)B + TILE - 1) / TILE, tw = (size_t)N * KW * TILE;
        std::vector<double> tmp(ntiles * tw, 0.0);
        std::vector<int> slot(B);
        CK(cudaMemcpy(slot.data(), h->d_slot_of, (size_t)B * 4, cudaMemcpyDeviceToHost));
        for (int b = 0; b < B; ++b)
            for (int k = 0; k < N; ++k)
                for (int f = 0; f < KW; ++f)
                    tmp[(size_t)(slot[b] / TILE) * tw + ((size_t)k * KW + f) * TILE + (slot[b] % TILE)] = src[((size_t)b * KW + f) * N + k];
        CK(cudaMemcpyAsync(h->kkt_tiles, tmp.data(), tmp.size() * 8, cudaMemcpyHostToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        return 0;
    }
    const size_t words = (field == MPCB200_F_SCAL) ? (size_t)cnt : (size_t)cnt * N;
    CK(cudaMemcpy2DAsync(h->ws + off, (size_t)h->L.stride * 8, src, words * 8, words * 8, (size_t)B, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int mpcb200_run_phase(mpcb200_handle* h, int phase, int B)
{
    int rc = check_batch(h, B);
    if (rc) return rc;
    CK(cudaSetDevice(h->device));
    h->spec = 0;  // single phases run the plain serial KKT attempts
    if (phase == MPCB200_PHASE_EVAL && (rc = launch_regroup(h, B))) return rc;
    if ((rc = launch_phase(h, phase, B, 0, 1, true))) return rc;
    CK(cudaStreamSynchronize(h->stream));
    ev_collect(h);
    return 0;
}

extern "C" int mpcb200_set_stream(mpcb200_handle* h, void* cuda_stream)
{
    if (!h) return MPCB200_E_INVALID;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
    return 0;
}

extern "C" int mpcb200_set_option(mpcb200_handle* h, int option, int value)
{
    if (!h) return MPCB200_E_INVALID;
    if (option == MPCB200_OPT_KKT_ATTEMPTS && value >= 0 && value <= 2) { h->spec_mode = value; return 0; }
    if (option == MPCB200_OPT_STREAM_REFILL_EVERY && value >= 1 && value <= 16) { h->refill_every = value; return 0; }
    return set_err(h, MPCB200_E_INVALID, "unknown option or value");
}

extern "C" int mpcb200_set_timing(mpcb200_handle* h, unsigned phase_mask)
{
    if (!h) return MPCB200_E_INVALID;
    h->timing_mask = phase_mask;
    return 0;
}

extern "C" int mpcb200_time_phase(mpcb200_handle* h, int phase, int B, int reps, int flush_l2, double* ms_per_launch)
{
    int rc = check_batch(h, B);
    if (rc) return rc;
    h->spec = 0;
    if (reps < 1) reps = 1;
    CK(cudaSetDevice(h->device));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    double total = 0.0;
    for (int r = 0; r < reps; ++r)
    {
        if (flush_l2) { flush_kernel<<<148 * 8, 256, 0, h->stream>>>(h->d_flush, h->flush_n); CK(cudaGetLastError()); }
        CK(cudaEventRecord(a, h->stream));
        if ((rc = launch_phase(h, phase, B, 0, 1, false))) return rc;
        CK(cudaEventRecord(b, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, a, b));
        total += ms;
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    if (ms_per_launch) *ms_per_launch = total / reps;
    return 0;
}

extern "C" int mpcb200_stats_get(const mpcb200_handle* hc, mpcb200_stats* out)
{
    mpcb200_handle* h = const_cast<mpcb200_handle*>(hc);
    if (!h || !out) return MPCB200_E_INVALID;
    unsigned long long cnt[2] = {0, 0};
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(cnt, h->d_counters, 16, cudaMemcpyDeviceToHost));
    h->stats.kkt_instances = (long long)cnt[0];
    h->stats.kkt_sweeps = (long long)cnt[1];
    *out = h->stats;
    return 0;
}
extern "C" int mpcb200_stats_reset(mpcb200_handle* h)
{
    if (!h) return MPCB200_E_INVALID;
    memset(&h->stats, 0, sizeof(h->stats));
    CK(cudaSetDevice(h->device));
    CK(cudaMemsetAsync(h->d_counters, 0, 16, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}
extern "C" int mpcb200_export_controls(mpcb200_handle* h, void* dst_dev)
{
    if (!h || h->B < 1 || !dst_dev) return set_err(h, MPCB200_E_INVALID, "nothing to export");
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(dst_dev, h->d_upacked, (size_t)h->B * (h->cfg.n - 1) * 16, cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}
extern "C" int mpcb200_costmap_obstacles(mpcb200_handle* h, int B, const mpcb200_costmaps* maps, const double* robot_pose, double behind_robot_dist,
                                         int max_per_instance, int* count, int* found, int* type, double* params)
{
    if (!h) return MPCB200_E_INVALID;
    if (B > 65535) return set_err(h, MPCB200_E_INVALID, "costmap_obstacles: at most 65535 robots per call");
    if (B < 1 || !maps || !maps->cost || !maps->origin || !robot_pose || !count || !type || !params || max_per_instance < 1)
        return set_err(h, MPCB200_E_INVALID, "costmap_obstacles: B >= 1, maps, poses and output arrays are required");
    if (maps->size_x < 2 || maps->size_y < 2 || !(maps->resolution > 0)) return set_err(h, MPCB200_E_INVALID, "costmap_obstacles: maps of at least 2 x 2 cells with a positive resolution");
    CK(cudaSetDevice(h->device));
    const size_t W = (size_t)maps->size_x, H = (size_t)maps->size_y, M = (size_t)max_per_instance;
    const int nrb = (int)((H - 1 + 31) / 32);            // 32-row blocks of the rows the reference visits
    const int Wp = (int)((W + 3) / 4 * 4);               // mask row pitch: four columns per marking thread
    const size_t mask_words = (size_t)B * nrb * Wp;
    const size_t need = (size_t)B * W * H + (size_t)B * 5 * 8 + 2 * (size_t)B * W * 4 + 2 * (size_t)B * 4 + (size_t)B * M * (MPCB200_OBST_STRIDE * 8 + 4) +
                        mask_words * 4 + 512;
    if (need > h->cm_cap)
    {
        if (h->d_cm) cudaFree(h->d_cm);
        h->d_cm = nullptr; h->cm_cap = 0;
        CK(cudaMalloc(&h->d_cm, need));
        h->cm_cap = need;
    }
    // carve the scratch: 16-byte aligned pieces first (mask rows are stored as uint4), then ints, then the maps
    char* p = (char*)h->d_cm;
    unsigned* d_mask = (unsigned*)p; p += mask_words * 4;
    double* d_params = (double*)p; p += (size_t)B * M * MPCB200_OBST_STRIDE * 8;
    double* d_origin = (double*)p; p += (size_t)B * 2 * 8;
    double* d_pose = (double*)p; p += (size_t)B * 3 * 8;
    int* d_colcount = (int*)p; p += (size_t)B * W * 4;
    int* d_colstart = (int*)p; p += (size_t)B * W * 4;
    int* d_count = (int*)p; p += (size_t)B * 4;
    int* d_found = (int*)p; p += (size_t)B * 4;
    int* d_type = (int*)p; p += (size_t)B * M * 4;
    unsigned char* d_cost = (unsigned char*)p;            // 4-byte aligned: everything before it is a multiple of 4 bytes
    CK(cudaMemcpyAsync(d_cost, maps->cost, (size_t)B * W * H, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d_origin, maps->origin, (size_t)B * 16, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d_pose, robot_pose, (size_t)B * 24, cudaMemcpyHostToDevice, h->stream));
    h->stats.h2d_bytes += (long long)((size_t)B * W * H + (size_t)B * 40);
    CostmapArgs a{maps->size_x, maps->size_y, maps->resolution, behind_robot_dist, d_cost, d_origin, d_pose};
    const dim3 grid_mark((unsigned)((W / 4 + 1 + 63) / 64), (unsigned)B), grid_emit((unsigned)((W + 127) / 128), (unsigned)B);
    cudaEvent_t t0, t1;
    CK(cudaEventCreate(&t0)); CK(cudaEventCreate(&t1));
    CK(cudaEventRecord(t0, h->stream));
    if (W % 4 == 0) costmap_mark_kernel<true><<<grid_mark, 64, 0, h->stream>>>(a, B, nrb, Wp, d_mask, d_colcount);
    else costmap_mark_kernel<false><<<grid_mark, 64, 0, h->stream>>>(a, B, nrb, Wp, d_mask, d_colcount);
    costmap_offsets_kernel<<<B, 256, 0, h->stream>>>(maps->size_x, B, d_colcount, d_colstart, max_per_instance, d_count, d_found);
    costmap_emit_kernel<<<grid_emit, 128, 0, h->stream>>>(a, B, nrb, Wp, d_mask, d_colstart, max_per_instance, d_params, d_type);
    CK(cudaGetLastError());
    CK(cudaEventRecord(t1, h->stream));
    h->stats.launches_total += 3;
    CK(cudaMemcpyAsync(count, d_count, (size_t)B * 4, cudaMemcpyDeviceToHost, h->stream));
    if (found) CK(cudaMemcpyAsync(found, d_found, (size_t)B * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(type, d_type, (size_t)B * M * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(params, d_params, (size_t)B * M * MPCB200_OBST_STRIDE * 8, cudaMemcpyDeviceToHost, h->stream));
    h->stats.d2h_bytes += (long long)((size_t)B * 8 + (size_t)B * M * (4 + MPCB200_OBST_STRIDE * 8));
    CK(cudaStreamSynchronize(h->stream));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, t0, t1));
    h->costmap_ms = ms;
    cudaEventDestroy(t0); cudaEventDestroy(t1);
    return MPCB200_OK;
}

extern "C" double mpcb200_costmap_last_ms(const mpcb200_handle* h) { return h ? h->costmap_ms : 0.0; }

extern "C" int mpcb200_flush_l2(mpcb200_handle* h)
{
    if (!h) return MPCB200_E_INVALID;
    CK(cudaSetDevice(h->device));
    flush_kernel<<<148 * 8, 256, 0, h->stream>>>(h->d_flush, h->flush_n);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}
