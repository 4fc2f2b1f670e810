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


// ---- footprint-vs-costmap feasibility of a pose trajectory: Controller::isPoseTrajectoryFeasible (src/controller.cpp:859-917,
//      called at src/mpc_local_planner_ros.cpp:414-428) for B robots.  One warp per robot, one lane per interval of the
//      trajectory: the lane checks pose i and the poses interpolated between i and i+1 (accumulated step by step, as the
//      reference does), every check rasterises the footprint edges over the robot's map (byte gathers).
//      [EXT] CostmapModel::footprintCost / lineCost / pointCost, LineIterator, Costmap2D::worldToMap restated from upstream
//      knowledge of ROS navigation (see oracle/mpc_oracle.c: orc_pose_trajectory_feasible). ----
struct FeasArgs
{
    int size_x, size_y;
    double resolution;
    const unsigned char* cost;   // [B][size_y][size_x]
    const double* origin;        // [B][2]
    const double* xseq;          // [B][n][3]
    int n;
    const double* footprint;     // [n_fp][2], robot frame
    int n_fp;
    double inscribed_radius, min_resolution_angular;
    int look_ahead_idx;
};
__device__ __forceinline__ bool feas_world_to_map(const FeasArgs& a, double ox, double oy, double wx, double wy, int* mx, int* my)
{
    if (wx < ox || wy < oy) return false;
    *mx = (int)((wx - ox) / a.resolution);
    *my = (int)((wy - oy) / a.resolution);
    return *mx < a.size_x && *my < a.size_y;
}
__device__ __forceinline__ double feas_line_cost(const FeasArgs& a, const unsigned char* map, int x0, int x1, int y0, int y1)
{
    const int deltax = abs(x1 - x0), deltay = abs(y1 - y0);
    int x = x0, y = y0, xinc1, xinc2, yinc1, yinc2, den, num, numadd, numpixels;
    if (x1 >= x0) { xinc1 = 1; xinc2 = 1; } else { xinc1 = -1; xinc2 = -1; }
    if (y1 >= y0) { yinc1 = 1; yinc2 = 1; } else { yinc1 = -1; yinc2 = -1; }
    if (deltax >= deltay) { xinc1 = 0; yinc2 = 0; den = deltax; num = deltax / 2; numadd = deltay; numpixels = deltax; }
    else { xinc2 = 0; yinc1 = 0; den = deltay; num = deltay / 2; numadd = deltax; numpixels = deltay; }
    double line_cost = 0.0;
    for (int cur = 0; cur <= numpixels; ++cur)
    {
        const unsigned char c = map[(size_t)y * a.size_x + x];
        const double pc = c == 255 ? -2.0 : (c == 254 ? -1.0 : (double)c);
        if (pc < 0) return pc;
        if (line_cost < pc) line_cost = pc;
        num += numadd;
        if (num >= den) { num -= den; x += xinc1; y += yinc1; }
        x += xinc2; y += yinc2;
    }
    return line_cost;
}
__device__ __forceinline__ double feas_footprint_cost(const FeasArgs& a, const unsigned char* map, double ox, double oy, double px, double py, double th)
{
    int cx, cy;
    if (!feas_world_to_map(a, ox, oy, px, py, &cx, &cy)) return -1.0;
    if (a.n_fp < 3)
    {
        const unsigned char c = map[(size_t)cy * a.size_x + cx];
        if (c == 255) return -2.0;
        if (c == 254 || c == 253) return -1.0;
        return (double)c;
    }
    double si, co;
    sincos(th, &si, &co);
    double fc = 0.0;
    for (int i = 0; i < a.n_fp; ++i)
    {
        const int j = (i + 1) % a.n_fp;
        const double fix = a.footprint[2 * i], fiy = a.footprint[2 * i + 1], fjx = a.footprint[2 * j], fjy = a.footprint[2 * j + 1];
        const double ax = px + (fix * co - fiy * si), ay = py + (fix * si + fiy * co);
        const double bx = px + (fjx * co - fjy * si), by = py + (fjx * si + fjy * co);
        int x0, y0, x1, y1;
        if (!feas_world_to_map(a, ox, oy, ax, ay, &x0, &y0)) return -3.0;
        if (!feas_world_to_map(a, ox, oy, bx, by, &x1, &y1)) return -3.0;
        const double lc = feas_line_cost(a, map, x0, x1, y0, y1);
        if (fc < lc) fc = lc;
        if (lc < 0) return lc;
    }
    return fc;
}
__global__ void feasible_kernel(FeasArgs a, int B, unsigned char* feasible)
{
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (b >= B) return;
    const unsigned char* map = a.cost + (size_t)b * a.size_x * a.size_y;
    const double ox = a.origin[2 * b], oy = a.origin[2 * b + 1];
    const double* xs = a.xseq + (size_t)b * a.n * 3;
    int look = a.look_ahead_idx;
    if (look < 0 || look >= a.n) look = a.n - 1;
    bool ok = a.n >= 2;
    for (int i = lane; i <= look && ok; i += 32)
    {
        const double px = xs[3 * i], py = xs[3 * i + 1], pth = xs[3 * i + 2];
        if (feas_footprint_cost(a, map, ox, oy, px, py, pth) == -1.0) { ok = false; break; }
        if (i < look)
        {
            const double delta_rot = normalize_theta(xs[3 * i + 5] - pth);
            const double dx = xs[3 * i + 3] - px, dy = xs[3 * i + 4] - py;
            const double dist = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
            if (fabs(delta_rot) > a.min_resolution_angular || dist > a.inscribed_radius)
            {
                const double ca = ceil(fabs(delta_rot) / a.min_resolution_angular), cb = ceil(dist / a.inscribed_radius);
                const int n_add = (int)(ca > cb ? ca : cb) - 1;
                double ix = px, iy = py, ith = pth;
                for (int step = 0; step < n_add; ++step)
                {
                    ix = ix + dx / (n_add + 1.0);
                    iy = iy + dy / (n_add + 1.0);
                    ith = normalize_theta(ith + delta_rot / (n_add + 1.0));
                    if (feas_footprint_cost(a, map, ox, oy, ix, iy, ith) == -1.0) { ok = false; break; }
                }
            }
        }
    }
    const bool all_ok = __all_sync(FULLMASK, ok) != 0;
    if (lane == 0) feasible[b] = all_ok ? 1 : 0;
}
