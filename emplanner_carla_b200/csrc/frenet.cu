// K4 - batched Cartesian <-> Frenet helpers, one warp per scene.
//
// Replaces reference planner/planning_utils.py
//   match_projection_points :364-426, cal_projection_s_fun :429-445, cal_s_map_fun :448-472,
//   cal_s_l_fun :475-509, cal_s_l_deri_fun :512-588
// and planner/path_planning.py
//   cal_proj_point :52-75, frenet_2_x_y_theta_kappa :15-46 (everything before the trailing smoother).
//
// The per-scene reference line ((x, y, theta, kappa) x n_nodes, 51 nodes = 1632 B in the reference,
// planning_utils.sampling :244-245) is staged into shared memory with a TMA bulk copy per CTA; this is the
// stage that actually consumes the reference line's heading and curvature.
// Bug-compatible details kept on purpose: Q12 every projection of a point list goes
// through the match index of the list's FIRST point (:413); the nearest-node scan stops after 50 consecutive
// non-improvements (:400); Q14/Q15 Frenet->Cartesian truncates at s_map[-1] (path_planning.py:40-41).
#include "emp_common.cuh"
#include "frenet_kernels.h"

namespace emp {
namespace {

// Nearest-node scan for one point: first minimum, early exit after `limit` consecutive non-improvements.
//   pre < 0 : match_projection_points / find_match_points(is_first_run=True): from node 0, limit 50 (:383-402)
//   pre >= 0: find_match_points(is_first_run=False): from node `pre`, forward if the point is ahead of that
//             node along its tangent, else backward, limit 5 (:117-167)
__device__ __forceinline__ int nearest_node(const double *line, int n_nodes, double x, double y, int pre) {
    double best = __longlong_as_double(0x7ff0000000000000LL);
    int start = 0, step = 1, limit = 50;
    if (pre >= 0) {
        start = min(pre, n_nodes - 1);
        limit = 5;
        const double th = line[4 * start + 2];
        const double flag = (x - line[4 * start]) * cos(th) + (y - line[4 * start + 1]) * sin(th);
        step = flag > 0 ? 1 : -1;
    }
    int best_i = 0, inc = 0;                       // match_point_index_list is zero-initialised (:69)
    for (int i = start; i >= 0 && i < n_nodes; i += step) {
        const double dx = line[4 * i] - x, dy = line[4 * i + 1] - y;
        const double d = sqrt(dx * dx + dy * dy);
        if (d < best) {
            best = d;
            best_i = i;
            inc = 0;
        } else if (++inc >= limit) {
            break;
        }
    }
    return best_i;
}

__global__ void __launch_bounds__(FR_WARPS * 32)
frenet_project_kernel(const double *__restrict__ lines, const double *__restrict__ origin,
                      const double *__restrict__ points, const int32_t *__restrict__ npts,
                      const double *__restrict__ vel, const double *__restrict__ acc,
                      const double *__restrict__ s_map_in, const int32_t *__restrict__ pre_match,
                      double *__restrict__ proj_out, double *__restrict__ s_map_out, int32_t *__restrict__ match_out, double *__restrict__ sl_out,
                      double *__restrict__ deriv_out, const FrenetLaunch p) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int scene0 = blockIdx.x * FR_WARPS;
    const int nsc = min(FR_WARPS, p.batch - scene0);
    const int N = p.n_nodes;
    const int line_words = 4 * N;
    double *slab = reinterpret_cast<double *>(smem + 16);
    double *smaps = slab + (size_t)FR_WARPS * line_words;
    const uint32_t bar = smem_u32(smem);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t bytes = (uint32_t)nsc * (uint32_t)line_words * 8u;
        mbar_expect_tx(bar, bytes);
        bulk_g2s(smem_u32(slab), lines + (size_t)scene0 * line_words, bytes, bar);
    }
    mbar_wait(bar, 0);
    if (warp >= nsc) return;
    const int sc = scene0 + warp;
    const double *line = slab + (size_t)warp * line_words;
    double *smap = smaps + (size_t)warp * N;

    const int pre = pre_match ? pre_match[sc] : -1;
    if (s_map_in != nullptr) {
        for (int i = lane; i < N; i += 32) smap[i] = s_map_in[(size_t)sc * N + i];
        __syncwarp();
    } else if (origin != nullptr) {
        // ---- cal_s_map_fun (:448-472): polyline length, origin projected through its nearest node
        for (int i = lane; i < N; i += 32) {
            double seg = 0.0;
            if (i > 0) {
                const double dx = line[4 * i] - line[4 * (i - 1)], dy = line[4 * i + 1] - line[4 * (i - 1) + 1];
                seg = sqrt(dx * dx + dy * dy);
            }
            smap[i] = seg;
        }
        __syncwarp();
        double s0 = 0.0;
        if (lane == 0) {
            double acc_s = 0.0;
            for (int i = 1; i < N; ++i) {                   // sequential, like ref_s_map.append(s + ref_s_map[-1])
                acc_s = smap[i] + acc_s;
                smap[i] = acc_s;
            }
            const double ox = origin[2 * sc], oy = origin[2 * sc + 1];
            const int mi = nearest_node(line, N, ox, oy, -1);
            const double th = line[4 * mi + 2];
            s0 = smap[mi] + ((ox - line[4 * mi]) * cos(th) + (oy - line[4 * mi + 1]) * sin(th));
        }
        s0 = shfl(s0, 0);
        __syncwarp();
        for (int i = lane; i < N; i += 32) {
            const double v = smap[i] - s0;
            smap[i] = v;
            if (s_map_out) s_map_out[(size_t)sc * N + i] = v;
        }
        __syncwarp();
    }
    if (points == nullptr) return;

    // ---- cal_s_l_fun (:475-509) / cal_s_l_deri_fun (:512-588): one lane per point
    const int P = p.max_points;
    const int np_ = min(npts ? npts[sc] : P, P);
    // Q12: every projection goes through the match index of the list's FIRST point (:413)
    int mi0 = 0;
    if (np_ > 0 && lane == 0) mi0 = nearest_node(line, N, points[(size_t)sc * P * 2], points[(size_t)sc * P * 2 + 1], pre);
    mi0 = __shfl_sync(EMP_FULL_MASK, mi0, 0);
    for (int base = 0; base < np_; base += 32) {
        const int k = base + lane;
        const bool on = k < np_;
        const double x = on ? points[((size_t)sc * P + k) * 2] : 0.0;
        const double y = on ? points[((size_t)sc * P + k) * 2 + 1] : 0.0;
        const int mi = on ? nearest_node(line, N, x, y, pre) : 0;
        if (on) {
            const double xm = line[4 * mi0], ym = line[4 * mi0 + 1], thm = line[4 * mi0 + 2], km = line[4 * mi0 + 3];
            const double ct = cos(thm), st = sin(thm);
            const double dsm = (x - xm) * ct + (y - ym) * st;
            const double xr = xm + dsm * ct, yr = ym + dsm * st;
            const double thr = thm + km * dsm;
            if (match_out) match_out[(size_t)sc * P + k] = mi;
            if (proj_out) {
                double *o = proj_out + ((size_t)sc * P + k) * 4;
                o[0] = xr; o[1] = yr; o[2] = thr; o[3] = km;
            }
            if (sl_out && (s_map_in != nullptr || origin != nullptr)) {
                const double tho = line[4 * mi + 2];
                const double s = smap[mi] + ((x - line[4 * mi]) * cos(tho) + (y - line[4 * mi + 1]) * sin(tho));
                const double l = (x - xr) * (-sin(thr)) + (y - yr) * cos(thr);
                sl_out[((size_t)sc * P + k) * 2] = s;
                sl_out[((size_t)sc * P + k) * 2 + 1] = l;
            }
            if (deriv_out && origin != nullptr) {
                // cal_s_l_deri_fun: l uses origin_xy for every point (:542), kappa' = 0 (:577)
                const double ox = origin[2 * sc], oy = origin[2 * sc + 1];
                const double nx = -sin(thr), ny = cos(thr), tx = cos(thr), ty = sin(thr);
                const double vx = vel[((size_t)sc * P + k) * 2], vy = vel[((size_t)sc * P + k) * 2 + 1];
                const double ax = acc[((size_t)sc * P + k) * 2], ay = acc[((size_t)sc * P + k) * 2 + 1];
                const double l = (ox - xr) * nx + (oy - yr) * ny;
                const double dl = vx * nx + vy * ny;
                const double dsd = (vx * tx + vy * ty) / (1 - km * l);
                const double ddl = (ax * nx + ay * ny) - km * (1 - km * l) * (dsd * dsd);
                const double l_ds = fabs(dsd) < 1e-6 ? 0.0 : dl / dsd;
                const double dds = ((ax * tx + ay * ty) + 2 * (dsd * dsd * km * l_ds) + dsd * dsd * 0.0 * l) / (1 - km * l);
                const double l_dds = fabs(dsd) < 1e-6 ? 0.0 : (ddl - l_ds * dds) / (dsd * dsd);
                double *o = deriv_out + ((size_t)sc * P + k) * 7;
                o[0] = l; o[1] = dl; o[2] = dsd; o[3] = ddl; o[4] = l_ds; o[5] = dds; o[6] = l_dds;
            }
        }
    }
}

__global__ void __launch_bounds__(FR_WARPS * 32)
frenet_to_cart_kernel(const double *__restrict__ lines, const double *__restrict__ s_maps,
                      const double *__restrict__ start_sl, const double *__restrict__ path_s,
                      const double *__restrict__ path_l, const int32_t *__restrict__ npath,
                      double *__restrict__ cart, int32_t *__restrict__ ncart, const FrenetLaunch p) {
    const int lane = threadIdx.x & 31;
    const int sc = blockIdx.x * FR_WARPS + (threadIdx.x >> 5);
    if (sc >= p.batch) return;
    const int N = p.n_nodes, M = p.max_points;
    const double *line = lines + (size_t)sc * 4 * N;
    const double *smap = s_maps + (size_t)sc * N;
    const int np_ = min(npath ? npath[sc] : M, M);
    const double s_last = smap[N - 1];
    double *out = cart + (size_t)sc * (M + 1) * 4;
    const double NaN = __longlong_as_double(0x7ff8000000000000LL);

    // cal_proj_point(plan_start_s, 0, ...) (:31): unguarded walk (Q15) -> flag -1 instead of IndexError
    const double ss = start_sl[2 * sc], sl0 = start_sl[2 * sc + 1];
    int idx0 = 0;
    bool oob = false;
    while (true) {
        if (idx0 + 1 >= N) { oob = true; break; }
        if (!(smap[idx0 + 1] < ss)) break;
        ++idx0;
    }
    if (oob) {
        if (lane == 0) ncart[sc] = -1;
        return;
    }
    if (lane == 0) {
        const double th = line[4 * idx0 + 2], kap = line[4 * idx0 + 3];
        const double ds = ss - smap[idx0];
        const double px = line[4 * idx0] + ds * cos(th), py = line[4 * idx0 + 1] + ds * sin(th);
        const double pth = th + kap * ds;
        out[0] = px + sl0 * (-sin(pth));
        out[1] = py + sl0 * cos(pth);
        out[2] = NaN;
        out[3] = NaN;
    }
    int running = idx0, count = 1;
    bool stop = false;
    for (int base = 0; base < np_ && !stop; base += 32) {
        const int k = base + lane;
        const bool on = k < np_;
        const double s = on ? path_s[(size_t)sc * M + k] : 0.0;
        const double l = on ? path_l[(size_t)sc * M + k] : 0.0;
        const bool beyond = on && (s > s_last);                       // truncation (:40-41)
        const unsigned bmask = __ballot_sync(EMP_FULL_MASK, beyond);
        const int first_bad = bmask ? __ffs(bmask) - 1 : 32;
        const bool use = on && lane < first_bad;
        int g = 0;
        if (use) {
            while (g + 1 < N - 1 && smap[g + 1] < s) ++g;             // s <= s_map[-1] keeps g+1 <= N-1
        }
        // walk is monotone from the previous index: idx_k = max(running, g_0..g_k)
        int m = use ? g : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(EMP_FULL_MASK, m, o);
            if (lane >= o) m = max(m, v);
        }
        m = max(m, running);
        if (use) {
            const double th = line[4 * m + 2], kap = line[4 * m + 3];
            const double ds = s - smap[m];
            const double px = line[4 * m] + ds * cos(th), py = line[4 * m + 1] + ds * sin(th);
            const double pth = th + kap * ds;
            double *o = out + (size_t)(1 + k) * 4;
            o[0] = px + l * (-sin(pth));
            o[1] = py + l * cos(pth);
            o[2] = pth;
            o[3] = kap;
        }
        const int n_used = min(first_bad, max(np_ - base, 0));
        running = __shfl_sync(EMP_FULL_MASK, m, max(min(n_used, 32) - 1, 0));
        count += min(n_used, 32);
        stop = first_bad < 32;
    }
    if (lane == 0) ncart[sc] = count;
}

}  // namespace

cudaError_t launch_frenet_project(const double *lines, const double *origin, const double *points,
                                  const int32_t *npts, const double *vel, const double *acc, const double *s_map_in,
                                  const int32_t *pre_match, double *proj, double *s_map, int32_t *match, double *sl,
                                  double *deriv, const FrenetLaunch &p, cudaStream_t st) {
    const int grid = (p.batch + FR_WARPS - 1) / FR_WARPS;
    const size_t smem = 16 + (size_t)FR_WARPS * (4 * p.n_nodes + p.n_nodes) * 8;
    cudaFuncSetAttribute(frenet_project_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    frenet_project_kernel<<<grid, FR_WARPS * 32, smem, st>>>(lines, origin, points, npts, vel, acc, s_map_in, pre_match,
                                                             proj, s_map, match, sl, deriv, p);
    return cudaGetLastError();
}

cudaError_t launch_frenet_to_cart(const double *lines, const double *s_maps, const double *start_sl,
                                  const double *path_s, const double *path_l, const int32_t *npath, double *cart,
                                  int32_t *ncart, const FrenetLaunch &p, cudaStream_t st) {
    const int grid = (p.batch + FR_WARPS - 1) / FR_WARPS;
    frenet_to_cart_kernel<<<grid, FR_WARPS * 32, 0, st>>>(lines, s_maps, start_sl, path_s, path_l, npath, cart, ncart, p);
    return cudaGetLastError();
}

}  // namespace emp
