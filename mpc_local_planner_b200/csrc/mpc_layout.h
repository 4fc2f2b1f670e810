// mpc_layout.h -- workspace layout of one OCP instance (shared by the CUDA library and the test-only CPU warp emulator)
#pragma once
#include "mpc_core.h"
#include "mpc_riccati_warp.h"

#define MAX_OBST 64          // obstacles resident per instance (the union of what the association selects along the horizon)
#define MAX_OBST_LIST 2048   // obstacles per instance in the caller's list (kept in global memory beyond MAX_OBST)
#define MAX_VP 8

HD inline bool kkt_is_ext(const Cfg& c) { return c.variable_dt || c.xf_fixed[0] || c.xf_fixed[1] || c.xf_fixed[2]; }

// The block of one instance is ONE contiguous run of doubles.  Its leading part, the RESIDENT PREFIX [0, oOBST + obstacles
// in use), is everything an interior-point iteration touches: the fused solve kernel keeps it in shared memory for the whole
// solve (the phase kernels stage it there per launch), so every offset below is valid for the shared-memory copy as well.
static inline void make_layout(const mpcb200_config* c, int M, int V, WsLayout& L)
{
    L.N = c->n; L.K = c->k_max_obstacles_per_stage; L.RS = 8 + L.K; L.M = M; L.V = V;
    const int N = L.N;
    int o = 0;
    auto take = [&](int n) { int r = o; o += (n + 1) & ~1; return r; };  // even offsets: every field is 16-byte aligned (bulk copies)
    // ---- state: what survives between two solves (warm start) and what the API reads back ----
    L.oSCAL = take(MPCB200_SCAL_WORDS);
    L.oIN = take(IN_WORDS);
    L.oX = take(3 * N); L.oU = take(2 * N); L.oNU = take(3 * N);
    L.oS = take(L.RS * N); L.oLAM = take(L.RS * N);
    L.oOBS = take(((L.K > 0 ? L.K : 1) * N + 7) / 8);   // one byte per (row slot, stage)
    L.oVPST = take(V > 0 ? V : 1);
    L.oVP = take((V > 0 ? V : 1) * 3);
    L.oSTEP = take(8 * N);
    L.oSTATE_END = o;   // end of the state: what a solve writes back to the instance block
    // ---- scratch of an iteration ----
    // DS (slack steps) and R0 (row residuals) live inside the line search only, the stage matrices of the KKT sweep only
    // inside the KKT phase (the evaluation parks the gradient parts in DS until its own end): they share their memory.
    {
        const int mm = kkt_is_ext(*c) ? rw_scratch_words<true>(N) : rw_scratch_words<false>(N);
        const int ls = 2 * ((L.RS * N + 1) & ~1);
        L.oMM = take(mm > ls ? mm : ls);
        L.oDS = L.oMM; L.oR0 = L.oMM + ((L.RS * N + 1) & ~1);
    }
    L.oOG = take(4 * (L.K > 0 ? L.K : 1) * N);
    L.oKKT = take(N * RSTR);                                                       // condensed KKT stage records [k][RSTR]
    // ---- inputs: the resident prefix ends after the obstacles in use (<= M) ----
    L.oOTYPE = take(M > 0 ? M : 1);
    L.oOBST = take((M > 0 ? M : 1) * MPCB200_OBST_STRIDE);
    // ---- global memory only ----
    L.oXINIT = take(3 * N);
    L.oOGIDX = take(M > 0 ? M : 1);   // list index of every resident obstacle (long lists; diagnostics and tests)
    L.stride = ((int64_t)o + 15) / 16 * 16;
}
// doubles of the resident prefix when m_used obstacles per instance are in use
static inline int resident_words(const WsLayout& L, int m_used) { return L.oOBST + MPCB200_OBST_STRIDE * ((m_used + 1) & ~1); }
