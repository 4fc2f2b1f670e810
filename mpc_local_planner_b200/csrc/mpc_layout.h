// mpc_layout.h -- workspace layout of one OCP instance (shared by the CUDA library and the test-only CPU warp emulator)
#pragma once
#include "mpc_core.h"
#include "mpc_riccati_lane.h"

#define MAX_OBST 64
#define MAX_VP 8

static inline void make_layout(const mpcb200_config* c, int M, int V, WsLayout& L)
{
    L.N = c->n; L.K = c->k_max_obstacles_per_stage; L.RS = 8 + L.K; L.M = M; L.V = V;
    const int N = L.N;
    int o = 0;
    auto take = [&](int n) { int r = o; o += (n + 1) & ~1; return r; };  // even offsets: every field is 16-byte aligned (bulk copies)
    // ---- instance image: what the eval / line-search kernels stage in shared memory, in this order ----
    L.oSCAL = take(MPCB200_SCAL_WORDS);
    L.oIN = take(IN_WORDS);
    L.oX = take(3 * N); L.oU = take(2 * N); L.oNU = take(3 * N);
    L.oS = take(L.RS * N); L.oLAM = take(L.RS * N);
    L.oOBS = take((L.K > 0 ? L.K : 1) * N);
    L.oVPST = take(V > 0 ? V : 1);
    L.oVP = take((V > 0 ? V : 1) * 3);
    L.oSTEP = take(8 * N);  /* (KKT records and Riccati gains live in 32-instance interleaved tiles, not here) */
    L.oDS = take(L.RS * N); L.oDLAM = take(L.RS * N);
    L.oOTYPE = take(M > 0 ? M : 1);
    L.oOBST = take((M > 0 ? M : 1) * MPCB200_OBST_STRIDE);  // image ends after the obstacles in use (<= M)
    // ---- global memory only ----
    L.oSTEP2 = take(8 * N);  /* step of the speculative second KKT attempt */
    L.oR0 = take(L.RS * N); L.oOG = take(4 * (L.K > 0 ? L.K : 1) * N);
    L.oXINIT = take(3 * N);
    L.stride = ((int64_t)o + 15) / 16 * 16;
}

