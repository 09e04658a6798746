// navgym_b200.cu — sm_100a kernels + C ABI for nav-gym's NavGym-v0 per-step hot path.
//
// Written from scratch for B200; the reference (leekwoon/nav-gym) has no GPU code.  What each
// piece replaces (paths relative to /root/reference/nav_gym/src/nav_gym_env/):
//   step_kernel          NavGymEnv.step            env.py:591-728 (robot branch)
//     kinematics         KetiRobot.set_vel         keti_robot.py:64-93
//     scan               NavGymEnv._compute_scan   env.py:385-441
//       march            range_libc calc_range_many (third party, call site env.py:425)
//       segments/discs   pymap2d render_contours_in_lidar / render_agents_in_lidar
//                                                  (third party, call sites env.py:430-432)
//     observation        _convert_obs              env.py:443-462
//     reward/done/info   compute_rewards/terminals/info   env.py:464-589
//     rollback           env.py:707-723
//   edt_*_kernel         range_libc PyOMap + PyRayMarching ctor   env.py:337-340
//
// Floating-point contract: compiled with -fmad=false; every arithmetic step that
// feeds an integer result (cells, hit cells, flags) is a single IEEE rounding, spelled with
// __f*_rn intrinsics where it matters, fmaf only where the contract says fused.
//
// One CTA = one environment, one thread = one lidar beam (512 threads).  The three scans a
// step may need (the step's scan, the crash re-scan env.py:718, the auto-reset first scan)
// run through ONE copy of the scan code inside a CTA-uniform pass loop.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/navgym_b200.h"

#define NB NAVGYM_NUM_BEAMS
#define OBS_DIM (NB + NAVGYM_OBS_TAIL)
#define HIT_NONE NAVGYM_HIT_NONE
#define EDT_INF_G 32768

static unsigned long long g_launches = 0;

// ------------------------------------------------------------------ small device helpers
__device__ __forceinline__ int xy_to_cell(float x32, double origin, double res, int dim, int rule)
{
    // batch_xy_to_ij, env.py:1235-1253 (rule 0: NumPy-1.x float64 division; 1: NumPy-2 float32)
    float c;
    if (rule == 0)
        c = (float)__ddiv_rn(__dsub_rn((double)x32, origin), res);
    else
        c = __fdiv_rn(__fsub_rn(x32, (float)origin), (float)res);
    if (c >= (float)dim) c = (float)(dim - 1);
    if (c < 0.0f) c = 0.0f;
    return __float2int_rz(c);
}

// range_libc RayMarching::calc_range, canonical form (oracle/navgym_oracle.c nvo_calc_range).
__device__ __forceinline__ float march(const float *__restrict__ dist, int W, int H, float x0,
                                       float y0, float dx, float dy, float max_range,
                                       float t_stop, int &hx, int &hy)
{
    float t = 0.0f;
    hx = HIT_NONE;
    hy = HIT_NONE;
    while (t < t_stop) {
        int px = __float2int_rz(__fmaf_rn(dx, t, x0));
        int py = __float2int_rz(__fmaf_rn(dy, t, y0));
        if ((unsigned)px >= (unsigned)W || (unsigned)py >= (unsigned)H) break;
        float d = __ldg(dist + (size_t)py * W + px);
        if (d <= 0.0f) {
            float xd = __fsub_rn((float)px, x0);
            float yd = __fsub_rn((float)py, y0);
            hx = (int)xd;
            hy = (int)yd;
            return __fsqrt_rn(__fadd_rn(__fmul_rn(xd, xd), __fmul_rn(yd, yd)));
        }
        t = __fadd_rn(t, fmaxf(__fmul_rn(d, 0.999f), 1.0f));
    }
    return max_range;
}

__device__ __forceinline__ float seg_hit(float ox, float oy, float dx, float dy, float ax,
                                         float ay, float bx, float by)
{
    float ex = __fsub_rn(bx, ax), ey = __fsub_rn(by, ay);
    float wx = __fsub_rn(ax, ox), wy = __fsub_rn(ay, oy);
    float den = __fsub_rn(__fmul_rn(dx, ey), __fmul_rn(dy, ex));
    if (den == 0.0f) return CUDART_INF_F;
    float tn = __fsub_rn(__fmul_rn(wx, ey), __fmul_rn(wy, ex));
    float un = __fsub_rn(__fmul_rn(wx, dy), __fmul_rn(wy, dx));
    float t = __fdiv_rn(tn, den);
    float u = __fdiv_rn(un, den);
    if (t >= 0.0f && u >= 0.0f && u <= 1.0f) return t;
    return CUDART_INF_F;
}

__device__ __forceinline__ float disc_hit(float ox, float oy, float dx, float dy, float X,
                                          float Y, float r)
{
    float cx = __fsub_rn(X, ox), cy = __fsub_rn(Y, oy);
    float b = __fadd_rn(__fmul_rn(dx, cx), __fmul_rn(dy, cy));
    float c = __fsub_rn(__fadd_rn(__fmul_rn(cx, cx), __fmul_rn(cy, cy)), __fmul_rn(r, r));
    float q = __fsub_rn(__fmul_rn(b, b), c);
    if (q < 0.0f) return CUDART_INF_F;
    float s = __fsqrt_rn(q);
    float t = __fsub_rn(b, s);
    if (t < 0.0f) t = __fadd_rn(b, s);
    if (t < 0.0f) return CUDART_INF_F;
    return t;
}

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k)
{
#pragma unroll
    for (int i = 0; i < 10; i++) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }

// standard normal for (env, episode, step, slot, beam): Philox4x32-10 + Box-Muller
__device__ __forceinline__ float beam_normal(uint64_t seed, uint32_t env, uint32_t episode,
                                             uint32_t step, uint32_t slot, uint32_t beam)
{
    uint4 r = philox4x32_10(make_uint4(env, episode, step, (slot << 16) | (beam >> 2)),
                            make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    uint32_t a = (beam & 2) ? r.z : r.x, b = (beam & 2) ? r.w : r.y;
    float rad = sqrtf(-2.0f * __logf(u01(a)));
    float s, c;
    __sincosf(6.283185307179586f * u01(b), &s, &c);
    return rad * ((beam & 1) ? s : c);
}

// ------------------------------------------------------------------ fused step kernel
struct __align__(16) StepSmem {
    int scan[NB];  // float bits; non-negative floats order like ints -> atomicMin works
    float dx[NB], dy[NB];
    float discs[NAVGYM_MAX_DISC * 3];
    float segs[NAVGYM_MAX_SEG * 4];
    double red[NB / 32];
    double px, py, th;         // pose the current pass scans from
    double ppx, ppy, pyaw;     // prev_obs fields
    double gx, gy, pv, pw;
    double c0, s0, c1, s1;     // cos/sin of theta before / after the turn
    double reward, dist;
    float lx, ly, lt;
    float res32, max_range;
    int ci, cj, W, H;
    int nd, ns;
    int map;
    int steps, episode;
    int crash, success, done, trunc;
    int next_pass;
    long long edt_off;
    float noise_std;
};

enum { PASS_STEP = 0, PASS_RESCAN = 1, PASS_RESET = 2, PASS_END = 3 };

// Angular window of beams that can see an obstacle spanning bearings [phi0, phi0 + width].
// Beam k looks along lin[k] + theta with lin[k] = ANGLE_MIN + k * step (env.py:388-390).
__device__ __forceinline__ void beam_window(float phi0, float width, float theta, int &k0, int &cnt)
{
    const float step = 0.012271843f, amin = -3.141592f;
    float rel = phi0 - theta - amin;
    rel -= 6.2831853f * floorf(rel * 0.15915494f);
    k0 = (int)floorf(rel / step) - 2;
    cnt = (int)ceilf(width / step) + 5;
    if (cnt > NB) cnt = NB;
}

// R rays of one thread marched in lockstep: the R dependent-load chains are independent of
// each other, so each iteration has R EDT gathers in flight per thread (latency hiding by
// ILP instead of by occupancy).  Per ray the arithmetic is exactly march()'s.
template <int R>
__device__ __forceinline__ void march_multi(const float *__restrict__ dist, int W, int H, float x0,
                                            float y0, const float (&dx)[R], const float (&dy)[R],
                                            float max_range, float t_stop, float (&rc)[R],
                                            int (&hx)[R], int (&hy)[R])
{
    float t[R];
    bool act[R];
#pragma unroll
    for (int j = 0; j < R; j++) {
        t[j] = 0.0f;
        act[j] = 0.0f < t_stop;
        rc[j] = max_range;
        hx[j] = HIT_NONE;
        hy[j] = HIT_NONE;
    }
    for (;;) {
        float d[R];
#pragma unroll
        for (int j = 0; j < R; j++) {
            d[j] = 1.0f;
            if (act[j]) {
                int px = __float2int_rz(__fmaf_rn(dx[j], t[j], x0));
                int py = __float2int_rz(__fmaf_rn(dy[j], t[j], y0));
                if ((unsigned)px >= (unsigned)W || (unsigned)py >= (unsigned)H) act[j] = false;
                else d[j] = __ldg(dist + (size_t)py * W + px);
            }
        }
        bool any = false;
#pragma unroll
        for (int j = 0; j < R; j++) {
            if (act[j]) {
                if (d[j] <= 0.0f) {
                    float xd = __fsub_rn((float)__float2int_rz(__fmaf_rn(dx[j], t[j], x0)), x0);
                    float yd = __fsub_rn((float)__float2int_rz(__fmaf_rn(dy[j], t[j], y0)), y0);
                    hx[j] = (int)xd;
                    hy[j] = (int)yd;
                    rc[j] = __fsqrt_rn(__fadd_rn(__fmul_rn(xd, xd), __fmul_rn(yd, yd)));
                    act[j] = false;
                } else {
                    t[j] = __fadd_rn(t[j], fmaxf(__fmul_rn(d[j], 0.999f), 1.0f));
                    act[j] = t[j] < t_stop;
                }
            }
            any |= act[j];
        }
        if (!any) break;
    }
}

// One CTA = one environment; TPB threads, each owning R = 512 / TPB beams (tid, tid + TPB, ...:
// a warp's lanes hold adjacent beams, whose gathers share cache lines).
template <bool IS_RESET_KERNEL, int TPB>
__global__ void __launch_bounds__(TPB, (TPB >= 512 ? 2 : (TPB == 256 ? 4 : 8)))
step_kernel(const navgym_step_args_t a)
{
    constexpr int R = NB / TPB;
    constexpr int NW = TPB / 32;
    __shared__ StepSmem sm;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int e = blockIdx.x;
    const int B = a.num_envs;
    double *S = a.state;
#define ST(f) S[(size_t)(f) * B + e]

    // ---------------- prologue: state, kinematics (keti_robot.py:64-93) -----------------
    if (warp == 0) {
        double th0 = ST(NAVGYM_S_TH);
        double v = 0, w = 0;
        if (!IS_RESET_KERNEL) {
            v = (double)a.actions[2 * e];
            w = (double)a.actions[2 * e + 1];
        }
        if (!IS_RESET_KERNEL && a.min_turn_radius > 0) {  // env.py:595-600
            double lim = __dmul_rn(fabs(w), a.min_turn_radius);
            if (v >= 0) v = v > lim ? v : lim;
            else v = v < -lim ? v : -lim;
        }
        double th1 = __dadd_rn(th0, __dmul_rn(w, a.dt));
        // lanes 0/1 evaluate the two sincos in parallel
        double sv, cv;
        sincos(lane == 0 ? th0 : th1, &sv, &cv);
        double s1 = __shfl_sync(0xffffffffu, sv, 1), c1 = __shfl_sync(0xffffffffu, cv, 1);
        if (lane == 0) {
            double px = ST(NAVGYM_S_PX), py = ST(NAVGYM_S_PY), thn = th0;
            sm.map = a.map_id[e];
            sm.steps = a.steps[e];
            sm.episode = a.episodes ? a.episodes[e] : 0;
            sm.noise_std = a.noise_std ? a.noise_std[e] : 0.0f;
            if (!IS_RESET_KERNEL) {
                double rx = __dadd_rn(__dmul_rn(0.14474, cv), px);
                double ry = __dadd_rn(__dmul_rn(0.14474, sv), py);
                rx = __dadd_rn(rx, __dmul_rn(__dmul_rn(c1, v), a.dt));
                ry = __dadd_rn(ry, __dmul_rn(__dmul_rn(s1, v), a.dt));
                px = __dadd_rn(__dmul_rn(-0.14474, c1), rx);
                py = __dadd_rn(__dmul_rn(-0.14474, s1), ry);
                const double twopi = 6.283185307179586;
                thn = fmod(th1, twopi);
                if (thn != 0 && thn < 0) thn = __dadd_rn(thn, twopi);
                sm.steps += 1;  // env.py:592
                sm.ppx = ST(NAVGYM_S_PPX); sm.ppy = ST(NAVGYM_S_PPY); sm.pyaw = ST(NAVGYM_S_PYAW);
                sm.pv = ST(NAVGYM_S_PV); sm.pw = ST(NAVGYM_S_PW);
                // prev_action after this step (env.py:725; the Ackermann clamp edits `action`)
                ST(NAVGYM_S_PV) = a.min_turn_radius > 0 ? v : (double)a.actions[2 * e];
                ST(NAVGYM_S_PW) = (double)a.actions[2 * e + 1];
            } else {
                sm.steps = 0;
                sm.ppx = px; sm.ppy = py; sm.pv = 0; sm.pw = 0; sm.pyaw = 0;
                ST(NAVGYM_S_PV) = 0; ST(NAVGYM_S_PW) = 0;
            }
            sm.px = px; sm.py = py; sm.th = thn;
            sm.gx = ST(NAVGYM_S_GX); sm.gy = ST(NAVGYM_S_GY);
            sm.next_pass = IS_RESET_KERNEL ? PASS_RESET : PASS_STEP;
            sm.crash = 0; sm.success = 0; sm.done = 0; sm.trunc = 0;
            sm.reward = 0; sm.dist = 0;
        }
    }
    // obstacles of this env -> shared (they do not move within a step)
    {
        int nd = a.discs ? min(a.ndisc[e], min(a.max_disc, NAVGYM_MAX_DISC)) : 0;
        int ns = a.segs ? min(a.nseg[e], min(a.max_seg, NAVGYM_MAX_SEG)) : 0;
        for (int i = tid; i < nd * 3; i += TPB) sm.discs[i] = a.discs[(size_t)e * a.max_disc * 3 + i];
        for (int i = tid; i < ns * 4; i += TPB) sm.segs[i] = a.segs[(size_t)e * a.max_seg * 4 + i];
        if (tid == 0) { sm.nd = nd; sm.ns = ns; }
    }
    float r[R];
    int pass = IS_RESET_KERNEL ? PASS_RESET : PASS_STEP;

    for (;;) {
        // ---- per-pass setup by thread 0: float32 lidar pose, origin cell (env.py:386,419)
        if (tid == 0) {
            const navgym_map_t m = a.maps[sm.map];
            sm.lx = (float)sm.px; sm.ly = (float)sm.py; sm.lt = (float)sm.th;
            sm.ci = xy_to_cell(sm.lx, m.ox, m.res, m.H, a.cell_rule);
            sm.cj = xy_to_cell(sm.ly, m.oy, m.res, m.W, a.cell_rule);
            sm.W = m.W; sm.H = m.H; sm.edt_off = m.edt_offset;
            sm.res32 = (float)m.res;
            sm.max_range = (float)((double)m.W * (double)m.H);
        }
        __syncthreads();
        // ---- beam directions + occupancy-grid march (env.py:388-390, 420-426)
        float dx[R], dy[R];
#pragma unroll
        for (int j = 0; j < R; j++) {
            const float h = (float)__dadd_rn(a.lin[tid + j * TPB], (double)sm.lt);
            double sd, cd;
            sincos((double)h, &sd, &cd);
            dx[j] = (float)cd;
            dy[j] = (float)sd;
        }
        {
            float rc[R];
            int hx[R], hy[R];
            march_multi<R>(a.edt_pool + sm.edt_off, sm.W, sm.H, (float)sm.ci, (float)sm.cj, dx, dy,
                           sm.max_range, fminf(a.t_stop, sm.max_range), rc, hx, hy);
            const bool rec = pass == (IS_RESET_KERNEL ? PASS_RESET : PASS_STEP) && a.hits;
#pragma unroll
            for (int j = 0; j < R; j++) {
                r[j] = __fmul_rn(rc[j], sm.res32);
                if (rec)
                    *reinterpret_cast<short2 *>(a.hits + ((size_t)e * NB + tid + j * TPB) * 2) =
                        make_short2((short)hx[j], (short)hy[j]);
            }
        }
        const int nobs = sm.nd + sm.ns;
        if (nobs > 0) {
            // ---- pedestrians: segments (env.py:430-431) and discs (env.py:432), min-merged.
            // One warp per obstacle, lanes across the beams of its angular window.
#pragma unroll
            for (int j = 0; j < R; j++) {
                sm.scan[tid + j * TPB] = __float_as_int(r[j]);
                sm.dx[tid + j * TPB] = dx[j];
                sm.dy[tid + j * TPB] = dy[j];
            }
            __syncthreads();
            const float ox = sm.lx, oy = sm.ly, th = sm.lt;
            for (int o = warp; o < nobs; o += NW) {
                int k0, cnt;
                if (o < sm.ns) {
                    const float ax = sm.segs[4 * o], ay = sm.segs[4 * o + 1];
                    const float bx = sm.segs[4 * o + 2], by = sm.segs[4 * o + 3];
                    float pa = atan2f(ay - oy, ax - ox), pb = atan2f(by - oy, bx - ox);
                    float d = pb - pa;
                    d -= 6.2831853f * rintf(d * 0.15915494f);
                    float da2 = (ax - ox) * (ax - ox) + (ay - oy) * (ay - oy);
                    float db2 = (bx - ox) * (bx - ox) + (by - oy) * (by - oy);
                    if (fabsf(d) > 3.0f || da2 < 1e-6f || db2 < 1e-6f) { k0 = 0; cnt = NB; }
                    else beam_window(d >= 0 ? pa : pb, fabsf(d), th, k0, cnt);
                    for (int i = lane; i < cnt; i += 32) {
                        int k = (k0 + i) & (NB - 1);
                        float t = seg_hit(ox, oy, sm.dx[k], sm.dy[k], ax, ay, bx, by);
                        if (t < CUDART_INF_F) atomicMin(&sm.scan[k], __float_as_int(t));
                    }
                } else {
                    const int q = o - sm.ns;
                    const float X = sm.discs[3 * q], Y = sm.discs[3 * q + 1], Rd = sm.discs[3 * q + 2];
                    float cx = X - ox, cy = Y - oy;
                    float dc = sqrtf(cx * cx + cy * cy);
                    if (dc <= Rd * 1.05f + 1e-3f) { k0 = 0; cnt = NB; }
                    else {
                        float half = asinf(fminf(Rd / dc, 1.0f)) * 1.01f + 1e-4f;
                        beam_window(atan2f(cy, cx) - half, 2.0f * half, th, k0, cnt);
                    }
                    for (int i = lane; i < cnt; i += 32) {
                        int k = (k0 + i) & (NB - 1);
                        float t = disc_hit(ox, oy, sm.dx[k], sm.dy[k], X, Y, Rd);
                        if (t < CUDART_INF_F) atomicMin(&sm.scan[k], __float_as_int(t));
                    }
                }
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < R; j++) r[j] = __int_as_float(sm.scan[tid + j * TPB]);
        }
        // ---- clip + noise (env.py:435-440)
        bool c_any = false, d_any = false;
#pragma unroll
        for (int j = 0; j < R; j++) {
            const int k = tid + j * TPB;
            float v = fminf(fmaxf(r[j], 0.0f), a.range_max);
            if (v != a.range_max) {
                if (a.noise) {
                    int slot = pass == PASS_RESCAN ? 1 : 0;
                    v = __fadd_rn(v, a.noise[((size_t)e * 2 + slot) * NB + k]);
                } else if (sm.noise_std > 0.0f) {
                    v = __fadd_rn(v, sm.noise_std * beam_normal(a.seed, (uint32_t)(a.env_offset + e),
                                                               (uint32_t)sm.episode, (uint32_t)sm.steps,
                                                               (uint32_t)pass, (uint32_t)k));
                }
            }
            r[j] = v;
            c_any |= v < a.thr[k];
            d_any |= v < a.dthr[k];
        }

        if (pass == PASS_STEP) {
            // ---- reward / done / info on this observation (env.py:464-589)
            const int crash = __syncthreads_or(c_any);
            const int discomf = __syncthreads_or(d_any) && !crash;
            if (discomf) {
                double q = CUDART_INF;
#pragma unroll
                for (int j = 0; j < R; j++) {
                    const int k = tid + j * TPB;
                    const float thr_k = a.thr[k], dthr_k = a.dthr[k];
                    float den = __fadd_rn(__fsub_rn(dthr_k, thr_k), 1e-6f);
                    q = fmin(q, __ddiv_rn(__dsub_rn((double)r[j], (double)thr_k), (double)den));
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) q = fmin(q, __shfl_xor_sync(0xffffffffu, q, o));
                if (lane == 0) sm.red[warp] = q;
                __syncthreads();
            }
            if (tid == 0) {
                double mn = CUDART_INF;
                if (discomf)
                    for (int i = 0; i < NW; i++) mn = fmin(mn, sm.red[i]);
                const double px = sm.px, py = sm.py;
                double dxg = __dsub_rn(sm.gx, px), dyg = __dsub_rn(sm.gy, py);
                double dist = sqrt(__dadd_rn(__dmul_rn(dxg, dxg), __dmul_rn(dyg, dyg)));
                double dxp = __dsub_rn(sm.gx, sm.ppx), dyp = __dsub_rn(sm.gy, sm.ppy);
                double pdist = sqrt(__dadd_rn(__dmul_rn(dxp, dxp), __dmul_rn(dyp, dyp)));
                int success = dist < a.dist_thresh;
                double r_s = success ? __dmul_rn(__dmul_rn(1.0, a.r_success), a.r_scale) : 0.0;
                double r_c = crash ? __dmul_rn(__dmul_rn(-1.0, a.r_crash), a.r_scale) : 0.0;
                double r_p = __dmul_rn(__dmul_rn(__dsub_rn(pdist, dist), a.r_progress), a.r_scale);
                double r_f = __dmul_rn(__dmul_rn(sm.pv, a.r_forward), a.r_scale);
                double r_r = __dmul_rn(__dmul_rn(__dmul_rn(-1.0, __dmul_rn(sm.pw, sm.pw)), a.r_rotation), a.r_scale);
                double r_d = 0.0;
                if (discomf)
                    r_d = __dmul_rn(__dmul_rn(-__dsub_rn(1.0, mn), a.r_discomfort), a.r_scale);
                double rew = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(r_s, r_c), r_p), r_f), r_r), r_d);
                int trunc = a.max_episode_steps > 0 && sm.steps >= a.max_episode_steps && !(success || crash);
                int done = success || crash || trunc;
                sm.crash = crash; sm.success = success; sm.done = done; sm.trunc = trunc;
                a.reward[e] = (float)rew;
                a.done[e] = (uint8_t)done;
                a.is_success[e] = (uint8_t)success;
                a.is_crash[e] = (uint8_t)crash;
                if (a.truncated) a.truncated[e] = (uint8_t)trunc;
                a.distance[e] = (float)dist;
                if (done && a.auto_reset) {
                    sm.next_pass = PASS_RESET;
                } else if (crash) {  // env.py:707-717: back to the pose / yaw of prev_obs
                    sm.px = sm.ppx; sm.py = sm.ppy; sm.th = sm.pyaw;
                    sm.next_pass = PASS_RESCAN;
                } else {
                    sm.next_pass = PASS_END;
                }
                if (sm.next_pass == PASS_RESET) {
                    // auto-reset: draw a spawn tuple (and a map) for the next episode
                    uint4 rnd = philox4x32_10(make_uint4((uint32_t)(a.env_offset + e), (uint32_t)sm.episode,
                                                         0x5eedu, 0xfffffff0u),
                                              make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32)));
                    int map = sm.map;
                    if (a.resample_map && a.num_maps > 1)
                        map = (int)(((uint64_t)rnd.y * (uint64_t)a.num_maps) >> 32);
                    const navgym_map_t m = a.maps[map];
                    if (m.spawn_count > 0) {
                        long long row = m.spawn_offset + (long long)(((uint64_t)rnd.x * (uint64_t)m.spawn_count) >> 32);
                        const double *sp = a.spawn_pool + row * 5;
                        sm.px = sp[0]; sm.py = sp[1]; sm.gx = sp[2]; sm.gy = sp[3]; sm.th = sp[4];
                        sm.map = map;
                    } else {  // no pool: restart from the rolled-back pose
                        sm.px = sm.ppx; sm.py = sm.ppy; sm.th = sm.pyaw;
                    }
                    sm.noise_std = a.noise_lo + (a.noise_hi - a.noise_lo) * u01(rnd.z);
                    sm.episode += 1;
                    sm.steps = 0;
                    sm.ppx = sm.px; sm.ppy = sm.py; sm.pv = 0; sm.pw = 0;
                    ST(NAVGYM_S_PV) = 0; ST(NAVGYM_S_PW) = 0;
                    ST(NAVGYM_S_GX) = sm.gx; ST(NAVGYM_S_GY) = sm.gy;
                    a.map_id[e] = sm.map;
                    if (a.noise_std) a.noise_std[e] = sm.noise_std;
                }
            }
            __syncthreads();
            pass = sm.next_pass;
            if (pass != PASS_END) continue;
        }
        break;
    }

    // ---------------- epilogue: observation row + state (env.py:455, 725-727) -----------
    float *o = a.obs + (size_t)e * a.obs_stride;
#pragma unroll
    for (int j = 0; j < R; j++) o[tid + j * TPB] = r[j];
    if (tid == 0) {
        double sn, cn;
        sincos(sm.th, &sn, &cn);
        double yaw = atan2(sn, cn);  // utils.py:5-9
        double t7[7] = {sm.ppx, sm.ppy, sm.px, sm.py, sm.pv, sm.pw, yaw};
#pragma unroll
        for (int i = 0; i < 7; i++) {
            o[NB + i] = (float)t7[i];
            if (a.tail64) a.tail64[(size_t)e * 7 + i] = t7[i];
        }
        ST(NAVGYM_S_PX) = sm.px; ST(NAVGYM_S_PY) = sm.py; ST(NAVGYM_S_TH) = sm.th;
        ST(NAVGYM_S_PPX) = sm.px; ST(NAVGYM_S_PPY) = sm.py; ST(NAVGYM_S_PYAW) = yaw;
        a.steps[e] = sm.steps;
        if (a.episodes) a.episodes[e] = sm.episode;
    }
#undef ST
}

// ------------------------------------------------------------------ EDT build kernels
// Pass 1: per column, distance to the nearest occupied cell of that column (coalesced in x).
__global__ void edt_columns_kernel(const uint8_t *__restrict__ occ, int H, int W, int32_t *__restrict__ g)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= W) return;
    int last = -1;
    for (int y = 0; y < H; y++) {
        if (occ[(size_t)y * W + x]) last = y;
        g[(size_t)y * W + x] = last < 0 ? EDT_INF_G : y - last;
    }
    last = -1;
    for (int y = H - 1; y >= 0; y--) {
        if (occ[(size_t)y * W + x]) last = y;
        int dn = last < 0 ? EDT_INF_G : last - y;
        size_t i = (size_t)y * W + x;
        if (dn < g[i]) g[i] = dn;
    }
}

// Pass 2: per row, exact integer minimisation d2(x) = min_q (x-q)^2 + g(q)^2 with the row in
// shared memory; the search window is |x-q| < g(x) (a farther q cannot beat q = x).
__global__ void edt_rows_kernel(const int32_t *__restrict__ g, int H, int W, float *__restrict__ dist)
{
    extern __shared__ int32_t row[];
    const int y = blockIdx.x;
    for (int x = threadIdx.x; x < W; x += blockDim.x) row[x] = g[(size_t)y * W + x];
    __syncthreads();
    for (int x = threadIdx.x; x < W; x += blockDim.x) {
        int gx = row[x];
        int best = gx * gx;  // <= 2^30
        int lo = max(0, x - gx + 1), hi = min(W - 1, x + gx - 1);
        for (int q = lo; q <= hi; q++) {
            int dq = x - q, gq = row[q];
            int v = dq * dq + gq * gq;
            best = min(best, v);
        }
        dist[(size_t)y * W + x] = __fsqrt_rn((float)best);
    }
}

// ------------------------------------------------------------------ stand-alone natives
__global__ void calc_range_many_kernel(const float *__restrict__ dist, int W, int H,
                                       const float *__restrict__ ins, float *__restrict__ outs, int N,
                                       float max_range, float t_stop, int16_t *__restrict__ hits)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float x0 = ins[3 * i], y0 = ins[3 * i + 1], h = ins[3 * i + 2];
    double s, c;
    sincos((double)h, &s, &c);
    int hx, hy;
    outs[i] = march(dist, W, H, x0, y0, (float)c, (float)s, max_range, t_stop, hx, hy);
    if (hits) { hits[2 * i] = (int16_t)hx; hits[2 * i + 1] = (int16_t)hy; }
}

__global__ void render_in_lidar_kernel(float *__restrict__ ranges, const float *__restrict__ headings,
                                       int K, const float *__restrict__ segs, int S,
                                       const float *__restrict__ discs, int D, float ox, float oy)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    double s, c;
    sincos((double)headings[k], &s, &c);
    float dx = (float)c, dy = (float)s, r = ranges[k];
    for (int i = 0; i < S; i++)
        r = fminf(r, seg_hit(ox, oy, dx, dy, segs[4 * i], segs[4 * i + 1], segs[4 * i + 2], segs[4 * i + 3]));
    for (int i = 0; i < D; i++)
        r = fminf(r, disc_hit(ox, oy, dx, dy, discs[3 * i], discs[3 * i + 1], discs[3 * i + 2]));
    ranges[k] = r;
}

// ------------------------------------------------------------------ C ABI
// threads per environment (rays per thread = 512 / TPB); NAVGYM_TPB overrides for tuning runs
static int g_tpb = 0;
static int step_tpb()
{
    if (g_tpb == 0) {
        const char *s = getenv("NAVGYM_TPB");
        int v = s ? atoi(s) : 128;
        g_tpb = (v == 512 || v == 256 || v == 128 || v == 64) ? v : 128;
    }
    return g_tpb;
}

template <bool RESET>
static void launch_step(const navgym_step_args_t &a, cudaStream_t st)
{
    switch (step_tpb()) {
    case 512: step_kernel<RESET, 512><<<a.num_envs, 512, 0, st>>>(a); break;
    case 256: step_kernel<RESET, 256><<<a.num_envs, 256, 0, st>>>(a); break;
    case 64: step_kernel<RESET, 64><<<a.num_envs, 64, 0, st>>>(a); break;
    default: step_kernel<RESET, 128><<<a.num_envs, 128, 0, st>>>(a); break;
    }
    g_launches++;
}

#define CK(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) return (int)_e; } while (0)

extern "C" {

int navgym_abi_version(void) { return 1; }
int navgym_sizeof_step_args(void) { return (int)sizeof(navgym_step_args_t); }
int navgym_sizeof_map(void) { return (int)sizeof(navgym_map_t); }
uint64_t navgym_launch_count(void) { return g_launches; }
const char *navgym_error_string(int code) { return cudaGetErrorString((cudaError_t)code); }
int navgym_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int navgym_step_batch(const navgym_step_args_t *args, void *stream)
{
    if (args->num_envs <= 0) return 0;
    if (args->obs_stride < OBS_DIM) return (int)cudaErrorInvalidValue;
    launch_step<false>(*args, (cudaStream_t)stream);
    return (int)cudaGetLastError();
}

int navgym_reset_obs_batch(const navgym_step_args_t *args, void *stream)
{
    if (args->num_envs <= 0) return 0;
    if (args->obs_stride < OBS_DIM) return (int)cudaErrorInvalidValue;
    launch_step<true>(*args, (cudaStream_t)stream);
    return (int)cudaGetLastError();
}

int navgym_edt_build(const uint8_t *occ_dev, int H, int W, float *dist_dev, int32_t *scratch_dev, void *stream)
{
    if (H <= 0 || W <= 0 || W > 12000) return (int)cudaErrorInvalidValue;
    cudaStream_t st = (cudaStream_t)stream;
    edt_columns_kernel<<<(W + 127) / 128, 128, 0, st>>>(occ_dev, H, W, scratch_dev);
    edt_rows_kernel<<<H, 256, (size_t)W * sizeof(int32_t), st>>>(scratch_dev, H, W, dist_dev);
    g_launches += 2;
    return (int)cudaGetLastError();
}

int navgym_calc_range_many(const float *dist_dev, int W, int H, const float *ins_dev, float *outs_dev,
                           int N, float max_range, float t_stop, int16_t *hits_dev, void *stream)
{
    if (N <= 0) return 0;
    calc_range_many_kernel<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        dist_dev, W, H, ins_dev, outs_dev, N, max_range, t_stop, hits_dev);
    g_launches++;
    return (int)cudaGetLastError();
}

int navgym_render_segments_in_lidar(float *ranges_dev, const float *headings_dev, int K,
                                    const float *segs_dev, int S, float ox, float oy, void *stream)
{
    if (K <= 0) return 0;
    render_in_lidar_kernel<<<(K + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        ranges_dev, headings_dev, K, segs_dev, S, nullptr, 0, ox, oy);
    g_launches++;
    return (int)cudaGetLastError();
}

int navgym_render_discs_in_lidar(float *ranges_dev, const float *headings_dev, int K,
                                 const float *discs_dev, int D, float ox, float oy, void *stream)
{
    if (K <= 0) return 0;
    render_in_lidar_kernel<<<(K + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        ranges_dev, headings_dev, K, nullptr, 0, discs_dev, D, ox, oy);
    g_launches++;
    return (int)cudaGetLastError();
}

int navgym_render_in_lidar_host(float *ranges_host, const float *headings_host, int K,
                                const float *segs_host, int S, const float *discs_host, int D,
                                float ox, float oy)
{
    if (K <= 0) return 0;
    float *buf = nullptr;
    size_t nf = (size_t)2 * K + (size_t)4 * S + (size_t)3 * D;
    CK(cudaMalloc(&buf, nf * sizeof(float)));
    float *r = buf, *hd = buf + K, *sg = hd + K, *dc = sg + 4 * S;
    cudaError_t err = cudaMemcpy(r, ranges_host, K * sizeof(float), cudaMemcpyHostToDevice);
    if (!err) err = cudaMemcpy(hd, headings_host, K * sizeof(float), cudaMemcpyHostToDevice);
    if (!err && S) err = cudaMemcpy(sg, segs_host, (size_t)4 * S * sizeof(float), cudaMemcpyHostToDevice);
    if (!err && D) err = cudaMemcpy(dc, discs_host, (size_t)3 * D * sizeof(float), cudaMemcpyHostToDevice);
    if (!err) {
        render_in_lidar_kernel<<<(K + 127) / 128, 128>>>(r, hd, K, sg, S, dc, D, ox, oy);
        g_launches++;
        err = cudaGetLastError();
    }
    if (!err) err = cudaMemcpy(ranges_host, r, K * sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(buf);
    return (int)err;
}

struct navgym_raymarching {
    int H, W;
    float max_range;
    float *dist;
};

navgym_raymarching_t *navgym_raymarching_create_host(const uint8_t *occ_host, int H, int W, float max_range)
{
    uint8_t *occ = nullptr;
    int32_t *scratch = nullptr;
    float *dist = nullptr;
    size_t n = (size_t)H * W;
    if (cudaMalloc(&occ, n) != cudaSuccess) return nullptr;
    if (cudaMalloc(&scratch, n * sizeof(int32_t)) != cudaSuccess) { cudaFree(occ); return nullptr; }
    if (cudaMalloc(&dist, n * sizeof(float)) != cudaSuccess) { cudaFree(occ); cudaFree(scratch); return nullptr; }
    cudaError_t err = cudaMemcpy(occ, occ_host, n, cudaMemcpyHostToDevice);
    if (!err) err = (cudaError_t)navgym_edt_build(occ, H, W, dist, scratch, nullptr);
    if (!err) err = cudaDeviceSynchronize();
    cudaFree(occ);
    cudaFree(scratch);
    if (err) { cudaFree(dist); return nullptr; }
    navgym_raymarching_t *rm = new navgym_raymarching_t{H, W, max_range, dist};
    return rm;
}

int navgym_raymarching_calc_range_many_host(navgym_raymarching_t *rm, const float *ins_host,
                                            float *outs_host, int N)
{
    if (!rm) return (int)cudaErrorInvalidValue;
    if (N <= 0) return 0;
    float *buf = nullptr;
    CK(cudaMalloc(&buf, (size_t)4 * N * sizeof(float)));
    cudaError_t err = cudaMemcpy(buf, ins_host, (size_t)3 * N * sizeof(float), cudaMemcpyHostToDevice);
    if (!err) err = (cudaError_t)navgym_calc_range_many(rm->dist, rm->W, rm->H, buf, buf + 3 * N, N,
                                                        rm->max_range, rm->max_range, nullptr, nullptr);
    if (!err) err = cudaMemcpy(outs_host, buf + 3 * N, (size_t)N * sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(buf);
    return (int)err;
}

// Reset-path host helper (no device work): 4-connected BFS distance over a grid, the
// uniform-cost equivalent of pyastar2d.astar_path (reference env.py:343-354).
void navgym_grid_bfs(const uint8_t *blocked, int H, int W, int sr, int sc, int32_t *dist)
{
    size_t n = (size_t)H * W;
    for (size_t i = 0; i < n; i++) dist[i] = -1;
    if (sr < 0 || sc < 0 || sr >= H || sc >= W || blocked[(size_t)sr * W + sc]) return;
    int32_t *queue = new int32_t[n];
    size_t head = 0, tail = 0;
    queue[tail++] = sr * W + sc;
    dist[(size_t)sr * W + sc] = 0;
    const int dr[4] = {1, -1, 0, 0}, dc[4] = {0, 0, 1, -1};
    while (head < tail) {
        int32_t cur = queue[head++];
        int r = cur / W, c = cur % W;
        for (int k = 0; k < 4; k++) {
            int nr = r + dr[k], nc = c + dc[k];
            if (nr < 0 || nc < 0 || nr >= H || nc >= W) continue;
            size_t ni = (size_t)nr * W + nc;
            if (blocked[ni] || dist[ni] >= 0) continue;
            dist[ni] = dist[cur] + 1;
            queue[tail++] = (int32_t)ni;
        }
    }
    delete[] queue;
}

const float *navgym_raymarching_edt_dev(const navgym_raymarching_t *rm) { return rm ? rm->dist : nullptr; }

void navgym_raymarching_destroy(navgym_raymarching_t *rm)
{
    if (!rm) return;
    cudaFree(rm->dist);
    delete rm;
}

}  // extern "C"
