// mpc_costmap.cuh -- costmap kernels of the C ABI (included by mpcb200.cu)
#pragma once
#include "mpc_device.cuh"

// ---- costmap -> point obstacles (MpcLocalPlannerROS::updateObstacleContainerWithCostmap, mpc_local_planner_ros.cpp:474-499) ----
// The reference walks the cells mx = 0..size_x-2 (outer), my = 0..size_y-2 (inner), keeps the LETHAL ones that are not farther
// than behind_dist behind the robot and appends them as point obstacles at the cell centres.  HBM-bound byte work, one byte per
// cell, read ONCE:
//   mark    a thread owns a tile of 16 adjacent columns x 32 rows (one 16-byte load per row, eight in flight), tests the words
//           for a LETHAL byte with one bit trick, applies the filter to the few hits and records them as one bit per cell in
//           per-column masks (32 rows per word, 1/8 byte per cell); tiles wholly behind the robot are not read at all;
//   emit    one CTA per robot: column counts (popcounts of the masks), their exclusive scan, then a thread per column walks
//           its mask words in row order: column offsets + bit order reproduce the reference's push_back order (mx outer, my
//           inner) exactly.
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
// the LETHAL cells of one column of a tile (bit r = row 32 rb + r) through the reference's "not far behind the robot" filter
__device__ __noinline__ unsigned costmap_filter_column(const CostmapArgs& a, unsigned bits, int mx, int rb, int b, double dirx, double diry)
{
    const double ox = a.origin[2 * b], oy = a.origin[2 * b + 1], px = a.pose[3 * b], py = a.pose[3 * b + 1];
    unsigned keep = 0u;
    while (bits)
    {
        const int r = __ffs(bits) - 1;
        bits &= bits - 1u;
        if (costmap_keep(a, mx, rb * 32 + r, ox, oy, px, py, dirx, diry)) keep |= 1u << r;
    }
    return keep;
}
// 0x80 in every byte of w that equals LETHAL (exact)
__device__ __forceinline__ unsigned lethal_bytes(unsigned w)
{
    const unsigned x = w ^ 0xFEFEFEFEu;
    return ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu);
}
// 16 cells of row `my` starting at column c0 as four words (cells beyond the row read as 0)
template <int VB>
__device__ __forceinline__ uint4 costmap_load16(const CostmapArgs& a, const unsigned char* map, int my, int c0)
{
    const unsigned char* q = map + (size_t)my * a.size_x + c0;
    if (VB == 16) return __ldg(reinterpret_cast<const uint4*>(q));
    unsigned v[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int g = 0; g < 4; ++g)
    {
        if (VB == 4) { if (c0 + 4 * g < a.size_x) v[g] = __ldg(reinterpret_cast<const unsigned*>(q + 4 * g)); }
        else
            for (int i = 0; i < 4; ++i)
                if (c0 + 4 * g + i < a.size_x) v[g] |= (unsigned)q[4 * g + i] << (8 * i);
    }
    return make_uint4(v[0], v[1], v[2], v[3]);
}
// Where a tile (columns c0..c1, rows r0..r1, cell centres) lies with respect to the reference's filter, decided on its corners
// with a safety margin far above the rounding of the per-cell test: 0 = every cell passes, 1 = every cell is rejected (behind the
// robot AND farther than behind_dist: such a tile is not even read), 2 = per-cell test needed.
__device__ __noinline__ int costmap_tile_class(const CostmapArgs& a, int b, double dirx, double diry, int c0, int c1, int r0, int r1)
{
    const double ox = a.origin[2 * b], oy = a.origin[2 * b + 1], px = a.pose[3 * b], py = a.pose[3 * b + 1];
    const double x0 = costmap_world(ox, c0, a.resolution) - px, x1 = costmap_world(ox, c1, a.resolution) - px;
    const double y0 = costmap_world(oy, r0, a.resolution) - py, y1 = costmap_world(oy, r1, a.resolution) - py;
    const double d00 = x0 * dirx + y0 * diry, d01 = x0 * dirx + y1 * diry, d10 = x1 * dirx + y0 * diry, d11 = x1 * dirx + y1 * diry;
    const double scale = fabs(x0) + fabs(x1) + fabs(y0) + fabs(y1) + fabs(px) + fabs(py) + fabs(ox) + fabs(oy) + 1.0;
    const double margin = 1e-9 * scale;
    const double dmin = fmin(fmin(d00, d01), fmin(d10, d11)), dmax = fmax(fmax(d00, d01), fmax(d10, d11));
    if (dmin > margin) return 0;                  // in front of the robot
    // nearest point of the tile rectangle to the robot
    const double nx = x0 > 0.0 ? x0 : (x1 < 0.0 ? x1 : 0.0), ny = y0 > 0.0 ? y0 : (y1 < 0.0 ? y1 : 0.0);
    const double fx = fmax(fabs(x0), fabs(x1)), fy = fmax(fabs(y0), fabs(y1));
    if (sqrt(fx * fx + fy * fy) < a.behind_dist - margin) return 0;   // inside the kept disc
    if (dmax < -margin && sqrt(nx * nx + ny * ny) > a.behind_dist + margin) return 1;
    return 2;
}
#define MARK_COLS 16
// A thread owns a tile of 16 columns x 32 rows.  Pass 1 streams the tile (one 16-byte load per row, eight in flight) and only
// notes which rows hold a LETHAL byte; pass 2 re-reads those few rows (L1/L2 hits) and sets the bits of the per-column masks;
// pass 3 (tiles the filter boundary crosses) applies the per-cell filter.  Keeping the rare work out of the streaming loop is
// what matters: with ~1 LETHAL cell per tile nearly every warp row has SOME lane with a hit.
template <int VB>   // bytes per load: 16 (size_x % 16 == 0), 4 (size_x % 4 == 0) or 1; the rows of every map start VB-aligned
__global__ void __launch_bounds__(256, 4) costmap_mark_kernel(CostmapArgs a, int B, int nrb, int ncg, int Wp, unsigned* mask /*[B][nrb][Wp]*/)
{
    // tiles of all robots in one index space: consecutive threads take consecutive 16-byte pieces of a row
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int tiles = ncg * nrb;
    const int b = (int)(t / tiles), tt = (int)(t - (long long)b * tiles);
    if (b >= B) return;
    const int rb = tt / ncg, c0 = MARK_COLS * (tt - rb * ncg);
    const unsigned char* map = a.cost + (size_t)b * a.size_x * a.size_y;
    const int rows = a.size_y - 1;
    uint4* dst = reinterpret_cast<uint4*>(mask + ((size_t)b * nrb + rb) * Wp + c0);
    const int c1 = min(c0 + MARK_COLS - 1, a.size_x - 2), r1 = min(rb * 32 + 31, rows - 1);
    double dirx, diry;
    sincos(a.pose[3 * b + 2], &diry, &dirx);   // PoseSE2::orientationUnitVec
    const int cls = c1 < c0 ? 1 : costmap_tile_class(a, b, dirx, diry, c0, c1, rb * 32, r1);
    if (cls == 1)
    {
#pragma unroll
        for (int g = 0; g < 4; ++g) dst[g] = make_uint4(0u, 0u, 0u, 0u);
        return;
    }
    unsigned hitrows = 0u;
#pragma unroll 1
    for (int r0 = 0; r0 < 32; r0 += 8)
    {
        uint4 w[8];
#pragma unroll
        for (int r = 0; r < 8; ++r)
        {
            const int my = rb * 32 + r0 + r;
            w[r] = make_uint4(0u, 0u, 0u, 0u);
            if (my < rows) w[r] = costmap_load16<VB>(a, map, my, c0);
        }
#pragma unroll
        for (int r = 0; r < 8; ++r)
        {
            const bool hit = word_has_lethal(w[r].x) | word_has_lethal(w[r].y) | word_has_lethal(w[r].z) | word_has_lethal(w[r].w);
            hitrows |= (hit ? 1u : 0u) << (r0 + r);   // (the quick test may flag a row without a LETHAL byte: pass 2 is exact)
        }
    }
    unsigned m[MARK_COLS];
#pragma unroll
    for (int i = 0; i < MARK_COLS; ++i) m[i] = 0u;
    while (hitrows)
    {
        const int r = __ffs(hitrows) - 1;
        hitrows &= hitrows - 1u;
        const uint4 w = costmap_load16<VB>(a, map, rb * 32 + r, c0);
        const unsigned z[4] = {lethal_bytes(w.x), lethal_bytes(w.y), lethal_bytes(w.z), lethal_bytes(w.w)};
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int i = 0; i < 4; ++i) m[4 * g + i] |= ((z[g] >> (8 * i + 7)) & 1u) << r;
    }
    // the reference never visits the last column
#pragma unroll
    for (int i = 0; i < MARK_COLS; ++i)
        if (c0 + i >= a.size_x - 1) m[i] = 0u;
    if (cls == 2)
    {
#pragma unroll
        for (int i = 0; i < MARK_COLS; ++i)
            if (m[i]) m[i] = costmap_filter_column(a, m[i], c0 + i, rb, b, dirx, diry);
    }
#pragma unroll
    for (int g = 0; g < 4; ++g) dst[g] = make_uint4(m[4 * g], m[4 * g + 1], m[4 * g + 2], m[4 * g + 3]);
}
// Column counts from the masks, their exclusive scan and the emission, for one robot (one CTA per robot): a thread owns a column,
// keeps its mask words in registers' reach (they are re-read from L1/L2), learns its start offset from the block scan and writes
// its cells in row order.  found = total, count = min(total, max_out).
__global__ void __launch_bounds__(256) costmap_emit_kernel(CostmapArgs a, int B, int nrb, int Wp, const unsigned* mask, int max_out, int* count, int* found,
                                                           double* params /*[B][max_out][MPCB200_OBST_STRIDE]*/, int* type /*[B][max_out]*/)
{
    const int b = blockIdx.x;
    __shared__ int warp_tot[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const double ox = a.origin[2 * b], oy = a.origin[2 * b + 1];
    int carry = 0;   // (every thread keeps its own copy: the totals are read by all)
    for (int base = 0; base < a.size_x; base += blockDim.x)
    {
        const int mx = base + threadIdx.x;
        int v = 0;
        if (mx < a.size_x - 1)
            for (int rb = 0; rb < nrb; ++rb) v += __popc(mask[((size_t)b * nrb + rb) * Wp + mx]);   // cells of column mx that become obstacles
        int incl = v;
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULLMASK, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) warp_tot[wid] = incl;
        __syncthreads();
        int woff = 0, tot = 0;
        for (int w = 0; w < nw; ++w) { if (w < wid) woff += warp_tot[w]; tot += warp_tot[w]; }
        __syncthreads();
        int o = carry + woff + incl - v;
        carry += tot;
        if (v == 0 || o >= max_out) continue;
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
