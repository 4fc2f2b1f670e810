// mpc_costmap.cuh -- costmap kernels of the C ABI (included by mpcb200.cu)
#pragma once
#include "mpc_device.cuh"

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

