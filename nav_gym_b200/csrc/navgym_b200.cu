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

// cos / sin of a beam heading, canonical form (DESIGN.md "beam direction"): Cody-Waite
// reduction by pi/2 in two fma steps, fdlibm kernel polynomials in Horner/fma form, quadrant
// fix-up.  A fixed sequence of IEEE operations, so the direction depends on the heading bits
// only (the CPU oracle evaluates the same sequence) — and ~5x fewer instructions than the
// full-range sincos() of the CUDA math library.
__device__ __forceinline__ void dir_sincos(double x, double &sn, double &cs)
{
    const double k = rint(__dmul_rn(x, 6.36619772367581382433e-01));
    double r = fma(-k, 1.57079632679489655800e+00, x);
    r = fma(-k, 6.12323399573676603587e-17, r);
    const double z = __dmul_rn(r, r);
    double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    ps = fma(z, ps, 2.75573137070700676789e-06);
    ps = fma(z, ps, -1.98412698298579493134e-04);
    ps = fma(z, ps, 8.33333333332248946124e-03);
    ps = fma(z, ps, -1.66666666666666324348e-01);
    const double s = fma(__dmul_rn(r, z), ps, r);
    double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    pc = fma(z, pc, -2.75573143513906633035e-07);
    pc = fma(z, pc, 2.48015872894767294178e-05);
    pc = fma(z, pc, -1.38888888888741095749e-03);
    pc = fma(z, pc, 4.16666666666666019037e-02);
    const double c = fma(__dmul_rn(z, z), pc, fma(z, -0.5, 1.0));
    const int n = __double2int_rn(k) & 3;
    const double a = (n & 1) ? c : s, b = (n & 1) ? s : c;
    sn = (n & 2) ? -a : a;
    cs = ((n + 1) & 2) ? -b : b;
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
// Optional per-phase cycle accounting (tools/phase_prof.py builds with -DNAVGYM_PROFILE).
#ifdef NAVGYM_PROFILE
__device__ unsigned long long g_prof[16];
#define PROF_DECL long long _pt = clock64(); unsigned long long _g_begin; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_g_begin));
#define PROF_MARK(i) do { if (tid == 0) { long long _n = clock64(); atomicAdd(&g_prof[i], (unsigned long long)(_n - _pt)); _pt = _n; } } while (0)
#else
#define PROF_DECL
#define PROF_MARK(i)
#endif

struct EnvSmem {
    int scan[NB];          // float bits of the ranges [m] (non-negative floats order like ints)
    float2 dir[NB];        // beam directions (cos, sin) of the current pass
    double red[16];
    // per-environment scalars parked here between the phases that need them, so the march
    // loop runs with a small register footprint
    double px, py, th, gx, gy, ppx, ppy, pyaw, pv, pw, act_v, act_w;
    double th_spec, yaw_spec;  // heading warp 1 assumed for the final pose, and its yaw
    int map, steps, episode, next_pass;
    int next_beam;         // next undealt entry of the survivor list
    int n_alive;           // beams still marching after the head phase
    short alive[NB];       // their indices
    float noise_std;
    // per-pass scan setup
    float lx, ly, lt, res32, max_range, t_stop;
    int ci, cj, W, H;
    long long edt_off;
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

// float -> cell index with C truncation semantics for x > -1, without the conversion pipe:
// 2^23 + x rounded toward zero leaves trunc(x) in the mantissa (x in [0, 2^23)).
__device__ __forceinline__ int trunc_cell(float x)
{
    return __float_as_int(__fadd_rz(fmaxf(x, 0.0f), 8388608.0f)) & 0x007fffff;
}

__device__ __forceinline__ void normal4(uint64_t seed, uint32_t env, uint32_t episode, uint32_t step,
                                        uint32_t slot, uint32_t group, float (&z)[4])
{
    uint4 r = philox4x32_10(make_uint4(env, episode, step, (slot << 16) | group),
                            make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    float r0 = sqrtf(-2.0f * __logf(u01(r.x))), r1 = sqrtf(-2.0f * __logf(u01(r.z)));
    float s0, c0, s1, c1;
    __sincosf(6.283185307179586f * u01(r.y), &s0, &c0);
    __sincosf(6.283185307179586f * u01(r.w), &s1, &c1);
    z[0] = r0 * c0; z[1] = r0 * s0; z[2] = r1 * c1; z[3] = r1 * s1;
}

// theta mod 2 pi with the sign of the divisor (numpy's float64 `%`, keti_robot.py:93).  One step
// turns by far less than 2 pi, so |x| < 4 pi in practice: there fmod is the identity or one
// exact subtraction (Sterbenz), bit-identical to the library call kept for anything larger.
__device__ __forceinline__ double wrap_2pi(double x)
{
    const double twopi = 6.283185307179586;
    double r;
    const double ax = fabs(x);
    if (ax < twopi) r = x;
    else if (ax < 2.0 * twopi) r = x < 0 ? __dadd_rn(x, twopi) : __dsub_rn(x, twopi);
    else r = fmod(x, twopi);
    if (r != 0 && r < 0) r = __dadd_rn(r, twopi);
    return r;
}

// Heading after this step's action (keti_robot.py:86-93); warps 0 and 1 both evaluate it.
__device__ __forceinline__ double turned_heading(double th0, double w, double dt, double &th1)
{
    th1 = __dadd_rn(th0, __dmul_rn(w, dt));
    return wrap_2pi(th1);
}

// Per-pass scan setup from the pose in shared memory: float32 lidar pose, origin cell
// (env.py:386, 419), map geometry.  Run by one thread.
__device__ __forceinline__ void pass_setup(EnvSmem &sm, const navgym_map_t &m, const navgym_step_args_t &a, int next_beam)
{
    sm.lx = (float)sm.px; sm.ly = (float)sm.py; sm.lt = (float)sm.th;
    sm.ci = xy_to_cell(sm.lx, m.ox, m.res, m.H, a.cell_rule);
    sm.cj = xy_to_cell(sm.ly, m.oy, m.res, m.W, a.cell_rule);
    sm.W = m.W; sm.H = m.H; sm.edt_off = m.edt_offset;
    sm.res32 = (float)m.res;
    sm.max_range = (float)((double)m.W * (double)m.H);
    sm.t_stop = fminf(fminf(a.t_stop, sm.max_range), 8.0e6f);
    sm.n_alive = 0;
    sm.next_beam = next_beam;
}

// Tail phase with S survivors per lane in flight, dealt from a shared counter (sm.next_beam
// starts at S * TPB): a slot whose beam ends takes the next undealt survivor.
template <int S, int TPB>
__device__ __forceinline__ void march_tail_slots(EnvSmem &sm, const float *__restrict__ dist, float x0, float y0,
                                                 int W, int H, float t_stop, int n_alive, int tid)
{
    const unsigned FULL = 0xffffffffu;
    int kb[S];  // the slot's current beam, -1 = none left
    float t[S], dx[S], dy[S];
#pragma unroll
    for (int s = 0; s < S; s++) {
        const int i = s * TPB + tid;
        kb[s] = i < n_alive ? (int)sm.alive[i] : -1;
        const int kk = kb[s] >= 0 ? kb[s] : 0;
        t[s] = __int_as_float(sm.scan[kk]);
        const float2 dd = sm.dir[kk];
        dx[s] = dd.x;
        dy[s] = dd.y;
    }
    if (n_alive <= 0) return;
    for (;;) {
        float d[S];
        int cx[S], cy[S];
        bool inb[S];
#pragma unroll
        for (int s = 0; s < S; s++) {
            cx[s] = __float2int_rz(__fmaf_rn(dx[s], t[s], x0));
            cy[s] = __float2int_rz(__fmaf_rn(dy[s], t[s], y0));
            inb[s] = ((unsigned)cx[s] < (unsigned)W) & ((unsigned)cy[s] < (unsigned)H);
            const unsigned idx = (inb[s] & (kb[s] >= 0)) ? (unsigned)(cy[s] * W + cx[s]) : 0u;
            d[s] = __ldg(dist + idx);
        }
#pragma unroll
        for (int s = 0; s < S; s++) {
            const bool hit = inb[s] & (d[s] <= 0.0f);
            float tn = __fadd_rn(t[s], fmaxf(__fmul_rn(d[s], 0.999f), 1.0f));
            const bool fin = !inb[s] | hit | !(tn < t_stop);
            if (fin & (kb[s] >= 0)) {
                // absolute hit cell, (y << 16 | x), or -1 for "no hit"
                sm.scan[kb[s]] = hit ? (cy[s] << 16 | cx[s]) : -1;
                const int i = atomicAdd(&sm.next_beam, 1);
                kb[s] = -1;
                if (i < n_alive) {
                    const int k = sm.alive[i];
                    kb[s] = k;
                    tn = __int_as_float(sm.scan[k]);
                    const float2 dd = sm.dir[k];
                    dx[s] = dd.x;
                    dy[s] = dd.y;
                }
            }
            t[s] = tn;
        }
        bool live = false;
#pragma unroll
        for (int s = 0; s < S; s++) live |= kb[s] >= 0;
        if (!__any_sync(FULL, live)) break;
    }
}

// One CTA = one environment, WPE warps.  Lane l of warp w owns beams l + 32 (w + WPE i),
// i = 0 .. 16/WPE - 1: at any moment the lanes of a warp work on neighbouring beams, whose
// EDT gathers share sectors.  Each lane walks its beams through MARCH_SLOTS independent march
// slots; a slot that finishes a beam immediately starts the lane's next one, so lanes stay
// busy instead of idling until the slowest beam of a lockstep group ends, and MARCH_SLOTS
// gathers per lane are in flight.  The three scans a step may need (the step's scan, the
// crash re-scan env.py:718, the auto-reset first scan) run through ONE copy of the scan code
// inside a CTA-uniform pass loop.
#ifndef NAVGYM_HEAD_STEPS
#define NAVGYM_HEAD_STEPS 4  // samples every beam marches in the lockstep head phase
#endif
#ifndef NAVGYM_RISK_MARGIN
#define NAVGYM_RISK_MARGIN 0.25f  // [m] clearance under which the next step may end the episode
#endif
#ifndef NAVGYM_THREADS_PER_SM
#define NAVGYM_THREADS_PER_SM 1024  // resident threads the register budget is tuned for
#endif
template <bool IS_RESET_KERNEL, int WPE, int MARCH_SLOTS>
__global__ void __launch_bounds__(WPE * 32, NAVGYM_THREADS_PER_SM / (WPE * 32)) step_kernel(const navgym_step_args_t a)
{
    constexpr int BPL = NB / (32 * WPE);  // beams per lane
    constexpr int TPB = WPE * 32;
    static_assert(BPL >= MARCH_SLOTS && BPL % MARCH_SLOTS == 0, "beams per lane vs slots");
    __shared__ EnvSmem sm;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int B = a.num_envs;
    const long long t_begin = clock64();
    // Which environment this CTA steps.  With a schedule buffer, CTAs take environments in
    // descending order of the cycles they cost in the previous step (they change slowly from
    // step to step), so the longest ones start first and the launch does not end on a lone
    // straggler; NAVGYM_SCHED_BUCKETS cost classes, bucket 0 = most expensive.
    int e = a.env_begin + blockIdx.x;
    int *sched_cnt = nullptr, *sched_list = nullptr;
    if (!IS_RESET_KERNEL && a.sched) {
        const int cur = a.sched_phase, nxt = (a.sched_phase + 1) % 3, clr = (a.sched_phase + 2) % 3;
        int *cnt = a.sched;                                   // [3][NBK]
        int *lst = a.sched + 3 * NAVGYM_SCHED_BUCKETS;        // [3][NBK][B]
        const int c = cnt[cur * NAVGYM_SCHED_BUCKETS + (threadIdx.x & 31)];
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if ((threadIdx.x & 31) >= o) incl += v;
        }
        const unsigned m = __ballot_sync(0xffffffffu, (int)blockIdx.x < incl);
        const int b = m ? __ffs(m) - 1 : 31;
        const int excl = __shfl_sync(0xffffffffu, incl - c, b);
        e = lst[((size_t)cur * NAVGYM_SCHED_BUCKETS + b) * B + ((int)blockIdx.x - excl)];
        sched_cnt = cnt + nxt * NAVGYM_SCHED_BUCKETS;
        sched_list = lst + (size_t)nxt * NAVGYM_SCHED_BUCKETS * B;
        if (blockIdx.x == 0 && threadIdx.x < NAVGYM_SCHED_BUCKETS) cnt[clr * NAVGYM_SCHED_BUCKETS + threadIdx.x] = 0;
    }
    const unsigned FULL = 0xffffffffu;
    double *S = a.state;
#define ST(f) S[(size_t)(f) * B + e]
#define BEAM(i) (tid + TPB * (i))

    PROF_DECL
    // ---------------- prologue (warp 0): state (lane f holds row f), kinematics ----------
    // Every global load the prologue needs is issued up front (they only depend on e), the map
    // descriptor as soon as the map id is back, so one L2 round trip overlaps the next and the
    // float64 kinematics.  Warp 1 meanwhile evaluates the yaw the epilogue will need (the same
    // heading arithmetic as warp 0): for every environment that neither rolls back nor resets,
    // the float64 sincos + atan2 of the final heading leave the critical path.
    if (WPE > 1 && warp == 1) {
        double th = ST(NAVGYM_S_TH);
        if (!IS_RESET_KERNEL) {
            double th1;
            th = turned_heading(th, (double)a.actions[2 * (size_t)e + 1], a.dt, th1);
        }
        double sn, cn;
        sincos(th, &sn, &cn);
        const double yaw = atan2(sn, cn);  // utils.py:5-9
        if (lane == 0) { sm.th_spec = th; sm.yaw_spec = yaw; }
    }
    if (warp == 0) {
        double sv = lane < NAVGYM_NS ? ST(lane) : 0.0;
        int steps = a.steps[e];
        const int map0 = a.map_id[e];
        const int episode0 = a.episodes ? a.episodes[e] : 0;
        const float noise_std0 = a.noise_std ? a.noise_std[e] : 0.0f;
        float2 av = make_float2(0.f, 0.f);
        if (!IS_RESET_KERNEL) av = *reinterpret_cast<const float2 *>(a.actions + 2 * (size_t)e);
        const navgym_map_t m0 = a.maps[map0];
        double px = __shfl_sync(FULL, sv, NAVGYM_S_PX), py = __shfl_sync(FULL, sv, NAVGYM_S_PY);
        double th0 = __shfl_sync(FULL, sv, NAVGYM_S_TH), th = th0;
        double act_v = 0, act_w = 0;
        if (!IS_RESET_KERNEL) {
            double v = (double)av.x, w = (double)av.y;
            if (a.min_turn_radius > 0) {  // env.py:595-600
                double lim = __dmul_rn(fabs(w), a.min_turn_radius);
                if (v >= 0) v = v > lim ? v : lim;
                else v = v < -lim ? v : -lim;
            }
            act_v = a.min_turn_radius > 0 ? v : (double)av.x;  // env.py:725 (the clamp edits `action`)
            act_w = (double)av.y;
            double th1;
            th = turned_heading(th0, w, a.dt, th1);
            double s_, c_;
            sincos(lane == 0 ? th0 : th1, &s_, &c_);  // lanes 0 / 1 in parallel
            double s0 = __shfl_sync(FULL, s_, 0), c0 = __shfl_sync(FULL, c_, 0);
            double s1 = __shfl_sync(FULL, s_, 1), c1 = __shfl_sync(FULL, c_, 1);
            // keti_robot.py:64-93
            double rx = __dadd_rn(__dmul_rn(0.14474, c0), px);
            double ry = __dadd_rn(__dmul_rn(0.14474, s0), py);
            rx = __dadd_rn(rx, __dmul_rn(__dmul_rn(c1, v), a.dt));
            ry = __dadd_rn(ry, __dmul_rn(__dmul_rn(s1, v), a.dt));
            px = __dadd_rn(__dmul_rn(-0.14474, c1), rx);
            py = __dadd_rn(__dmul_rn(-0.14474, s1), ry);
            steps += 1;  // env.py:592
        } else {
            steps = 0;
        }
        if (lane == NAVGYM_S_GX) sm.gx = sv;
        if (lane == NAVGYM_S_GY) sm.gy = sv;
        if (!IS_RESET_KERNEL) {
            if (lane == NAVGYM_S_PPX) sm.ppx = sv;
            if (lane == NAVGYM_S_PPY) sm.ppy = sv;
            if (lane == NAVGYM_S_PYAW) sm.pyaw = sv;
            if (lane == NAVGYM_S_PV) sm.pv = sv;
            if (lane == NAVGYM_S_PW) sm.pw = sv;
        }
        if (lane == 0) {
            sm.px = px; sm.py = py; sm.th = th;
            sm.act_v = act_v; sm.act_w = act_w;
            sm.map = map0;
            sm.steps = steps;
            sm.episode = episode0;
            sm.noise_std = noise_std0;
            if (IS_RESET_KERNEL) { sm.ppx = px; sm.ppy = py; sm.pyaw = 0; sm.pv = 0; sm.pw = 0; }
            if (WPE == 1) sm.th_spec = CUDART_NAN;
            pass_setup(sm, m0, a, TPB * MARCH_SLOTS);  // first pass: the map descriptor is already here
        }
    }
    int nd = a.discs ? min(a.ndisc[e], a.max_disc) : 0;
    int ns = a.segs ? min(a.nseg[e], a.max_seg) : 0;
    int pass = IS_RESET_KERNEL ? PASS_RESET : PASS_STEP;
    float *orow = a.obs + (size_t)e * a.obs_stride;

    long long t_pass = t_begin;  // start of the pass that produces the returned observation
    float margin = CUDART_INF_F; // its smallest clearance over the crash thresholds [m]
    for (bool first = true;; first = false) {
        // ---- per-pass setup; the first pass was set up by the prologue
        if (WPE > 1) __syncthreads(); else __syncwarp();
        if (!first) {
            t_pass = clock64();
            if (tid == 0) pass_setup(sm, a.maps[sm.map], a, TPB * MARCH_SLOTS);
            if (WPE > 1) __syncthreads(); else __syncwarp();
        }
        margin = CUDART_INF_F;
        PROF_MARK(0);
        const float lx = sm.lx, ly = sm.ly, lt = sm.lt;
        PROF_MARK(1);
        // ---- occupancy-grid march (env.py:425-426).
        // Every beam of a scan starts on the origin cell, so that first sample (t = 0) is taken
        // once per environment: either the origin is occupied (all beams end there) or all
        // beams advance by the same first step.  The march then runs in two phases:
        //  head: every thread marches its own beams NAVGYM_HEAD_STEPS samples, HB beams at a
        //        time in lockstep — almost every beam is still alive that early, and the HB
        //        independent EDT gathers per thread hide the L2 latency by ILP;
        //  tail: the surviving beams are compacted into a list and dealt out dynamically (a
        //        lane takes the next survivor whenever its beam ends), so lanes stay busy on
        //        the long-tailed remainder instead of idling until the slowest beam ends.
        // The loops only find each beam's hit cell (packed into sm.scan); ranges are computed
        // afterwards with all lanes active.
        {
            const int W = sm.W, H = sm.H, ci = sm.ci, cj = sm.cj;
            const float x0 = (float)ci, y0 = (float)cj;
            const float t_stop = sm.t_stop;
            const float *dist = a.edt_pool + sm.edt_off;
            asm volatile("" : "+l"(dist));  // keep base + offset folded into one register pair
            const float d0 = __ldg(dist + cj * W + ci);   // origin cell is clipped into the map
            const float t1 = fmaxf(__fmul_rn(d0, 0.999f), 1.0f);  // == 0.0f + first step
            const bool degenerate = (d0 <= 0.0f) | !(t1 < t_stop);
            constexpr int HB = BPL >= 4 ? 4 : BPL;   // beams in flight per thread in the head phase
#pragma unroll 1
            for (int r = 0; r < BPL / HB; r++) {
                float th_[HB], dxh[HB], dyh[HB];
                // beam directions (env.py:388-390, 420-424)
#pragma unroll
                for (int j = 0; j < HB; j++) {
                    const int k = BEAM(r * HB + j);
                    const float h = (float)__dadd_rn(a.lin[k], (double)lt);
                    double sd, cd;
                    dir_sincos((double)h, sd, cd);
                    dxh[j] = (float)cd;
                    dyh[j] = (float)sd;
                    sm.dir[k] = make_float2(dxh[j], dyh[j]);
                    th_[j] = t1;
                    if (degenerate) { sm.scan[k] = d0 <= 0.0f ? (cj << 16 | ci) : -1; th_[j] = -1.0f; }
                }
#pragma unroll 1
                for (int st = 0; st < NAVGYM_HEAD_STEPS; st++) {
                    float dv[HB];
                    int cx[HB], cy[HB];
                    bool inb[HB];
#pragma unroll
                    for (int j = 0; j < HB; j++) {
                        cx[j] = __float2int_rz(__fmaf_rn(dxh[j], th_[j], x0));
                        cy[j] = __float2int_rz(__fmaf_rn(dyh[j], th_[j], y0));
                        inb[j] = ((unsigned)cx[j] < (unsigned)W) & ((unsigned)cy[j] < (unsigned)H);
                        const unsigned idx = (inb[j] & (th_[j] >= 0.0f)) ? (unsigned)(cy[j] * W + cx[j]) : 0u;
                        dv[j] = __ldg(dist + idx);
                    }
#pragma unroll
                    for (int j = 0; j < HB; j++) {
                        const bool alive = th_[j] >= 0.0f;
                        const bool hit = inb[j] & (dv[j] <= 0.0f);
                        const float tn = __fadd_rn(th_[j], fmaxf(__fmul_rn(dv[j], 0.999f), 1.0f));
                        const bool fin = !inb[j] | hit | !(tn < t_stop);
                        if (alive & fin) sm.scan[BEAM(r * HB + j)] = hit ? (cy[j] << 16 | cx[j]) : -1;
                        th_[j] = (alive & !fin) ? tn : -1.0f;
                    }
                }
                // survivors: park t in the scan slot and append the beam to the compact list
#pragma unroll
                for (int j = 0; j < HB; j++) {
                    const int k = BEAM(r * HB + j);
                    const bool alive = th_[j] >= 0.0f;
                    if (alive) sm.scan[k] = __float_as_int(th_[j]);
                    const unsigned mk = __ballot_sync(FULL, alive);
                    int base = 0;
                    if (lane == 0 && mk) base = atomicAdd(&sm.n_alive, __popc(mk));
                    base = __shfl_sync(FULL, base, 0);
                    if (alive) sm.alive[base + __popc(mk & ((1u << lane) - 1u))] = (short)k;
                }
            }
            if (WPE > 1) __syncthreads(); else __syncwarp();  // any lane may be dealt any survivor
            PROF_MARK(2);
            {
                const int n_alive = sm.n_alive;
                if (MARCH_SLOTS == 1) {
                    // Warp w owns list entries w, w + WPE, w + 2 WPE, ...; they are dealt to its
                    // lanes with ballot ranks (no atomics, no cross-warp traffic): a lane whose
                    // beam ends takes the warp's next undealt entry.
                    int next_j = 32;                       // warp-uniform: entries dealt so far
                    int idx = warp + WPE * lane;
                    int kb = idx < n_alive ? (int)sm.alive[idx] : -1;
                    float t = __int_as_float(sm.scan[kb >= 0 ? kb : 0]);
                    float2 dd = sm.dir[kb >= 0 ? kb : 0];
                    if (__any_sync(FULL, kb >= 0)) {
                        for (;;) {
                            const int cx = __float2int_rz(__fmaf_rn(dd.x, t, x0));
                            const int cy = __float2int_rz(__fmaf_rn(dd.y, t, y0));
                            const bool inb = ((unsigned)cx < (unsigned)W) & ((unsigned)cy < (unsigned)H);
                            const unsigned ci_ = (inb & (kb >= 0)) ? (unsigned)(cy * W + cx) : 0u;
                            const float d = __ldg(dist + ci_);
                            const bool hit = inb & (d <= 0.0f);
                            t = __fadd_rn(t, fmaxf(__fmul_rn(d, 0.999f), 1.0f));
                            const bool fin = (kb >= 0) & (!inb | hit | !(t < t_stop));
                            const unsigned fm = __ballot_sync(FULL, fin);
                            if (fin) {
                                // absolute hit cell, (y << 16 | x), or -1 for "no hit"
                                sm.scan[kb] = hit ? (cy << 16 | cx) : -1;
                                idx = warp + WPE * (next_j + __popc(fm & ((1u << lane) - 1u)));
                                kb = -1;
                                if (idx < n_alive) {
                                    kb = sm.alive[idx];
                                    t = __int_as_float(sm.scan[kb]);
                                    dd = sm.dir[kb];
                                }
                            }
                            next_j += __popc(fm);
                            if (!__any_sync(FULL, kb >= 0)) break;
                        }
                    }
                } else {
                    march_tail_slots<MARCH_SLOTS, TPB>(sm, dist, x0, y0, W, H, t_stop, n_alive, tid);
                }
            }
            if (WPE > 1) __syncthreads(); else __syncwarp();
            // ranges (env.py:426), all lanes active: sqrt(di^2 + dj^2) * resolution
            const bool rec = pass == (IS_RESET_KERNEL ? PASS_RESET : PASS_STEP) && a.hits;
            const float max_range = sm.max_range, res32 = sm.res32;
#pragma unroll
            for (int i = 0; i < BPL; i++) {
                const int k = BEAM(i);
                const int cell = sm.scan[k];
                float rc = max_range;
                int rel = (int)0x80008000;
                if (cell != -1) {
                    const int hx = (cell & 0xffff) - ci, hy = (cell >> 16) - cj;
                    const float xd = (float)hx, yd = (float)hy;
                    rc = __fsqrt_rn(__fadd_rn(__fmul_rn(xd, xd), __fmul_rn(yd, yd)));
                    rel = (hx & 0xffff) | (hy << 16);
                }
                sm.scan[k] = __float_as_int(__fmul_rn(rc, res32));
                if (rec) reinterpret_cast<int *>(a.hits)[(size_t)e * NB + k] = rel;
            }
        }
        PROF_MARK(3);
        // ---- pedestrians: segments (env.py:430-431) and discs (env.py:432), min-merged.
        // One warp per obstacle, lanes across the beams of its angular window.
        if (ns + nd > 0) {
            if (WPE > 1) __syncthreads(); else __syncwarp();
            // (the first scan after an auto-reset sees the next episode's pedestrians, if given)
            const bool nxt = !IS_RESET_KERNEL && pass == PASS_RESET && a.discs_reset != nullptr;
            const float *discs = (nxt ? a.discs_reset : a.discs) + (size_t)e * a.max_disc * 3;
            const float *segs = (nxt ? a.segs_reset : a.segs) + (size_t)e * a.max_seg * 4;
            for (int o = warp; o < ns + nd; o += WPE) {
                int k0, cnt;
                if (o < ns) {
                    const float4 sg = *reinterpret_cast<const float4 *>(segs + 4 * o);
                    const float ax = sg.x, ay = sg.y, bx = sg.z, by = sg.w;
                    float pa = atan2f(ay - ly, ax - lx), pb = atan2f(by - ly, bx - lx);
                    float dl = pb - pa;
                    dl -= 6.2831853f * rintf(dl * 0.15915494f);
                    float da2 = (ax - lx) * (ax - lx) + (ay - ly) * (ay - ly);
                    float db2 = (bx - lx) * (bx - lx) + (by - ly) * (by - ly);
                    if (fabsf(dl) > 3.0f || da2 < 1e-6f || db2 < 1e-6f) { k0 = 0; cnt = NB; }
                    else beam_window(dl >= 0 ? pa : pb, fabsf(dl), lt, k0, cnt);
                    for (int i = lane; i < cnt; i += 32) {
                        const int k = (k0 + i) & (NB - 1);
                        const float tt = seg_hit(lx, ly, sm.dir[k].x, sm.dir[k].y, ax, ay, bx, by);
                        if (tt < CUDART_INF_F) atomicMin(&sm.scan[k], __float_as_int(tt));
                    }
                } else {
                    const int q = o - ns;
                    const float X = discs[3 * q], Y = discs[3 * q + 1], Rd = discs[3 * q + 2];
                    const float cx = X - lx, cy = Y - ly;
                    const float dc = sqrtf(cx * cx + cy * cy);
                    if (dc <= Rd * 1.05f + 1e-3f) { k0 = 0; cnt = NB; }
                    else {
                        const float half = asinf(fminf(Rd / dc, 1.0f)) * 1.01f + 1e-4f;
                        beam_window(atan2f(cy, cx) - half, 2.0f * half, lt, k0, cnt);
                    }
                    for (int i = lane; i < cnt; i += 32) {
                        const int k = (k0 + i) & (NB - 1);
                        const float tt = disc_hit(lx, ly, sm.dir[k].x, sm.dir[k].y, X, Y, Rd);
                        if (tt < CUDART_INF_F) atomicMin(&sm.scan[k], __float_as_int(tt));
                    }
                }
            }
            if (WPE > 1) __syncthreads(); else __syncwarp();
        }
        PROF_MARK(4);
        // ---- clip + noise (env.py:435-440), thresholds, observation row
        bool c_any = false, d_any = false;
        {
            const int nslot = pass == PASS_RESCAN ? 1 : 0;
            const float noise_std = sm.noise_std;
            const int steps = sm.steps, episode = sm.episode;
            const int SS = a.num_scan_stack > 1 ? a.num_scan_stack : 1;
            constexpr int G = BPL >= 4 ? 4 : BPL;
#pragma unroll 1
            for (int g = 0; g < BPL / G; g++) {
                float z[4] = {0.f, 0.f, 0.f, 0.f};
                if (!a.noise && noise_std > 0.0f)
                    normal4(a.seed, (uint32_t)(a.env_offset + e), (uint32_t)episode, (uint32_t)steps,
                            (uint32_t)pass, (uint32_t)(tid + TPB * g), z);
#pragma unroll
                for (int j = 0; j < G; j++) {
                    const int k = BEAM(G * g + j);
                    float v = fminf(fmaxf(__int_as_float(sm.scan[k]), 0.0f), a.range_max);
                    if (v != a.range_max) {
                        if (a.noise) v = __fadd_rn(v, a.noise[((size_t)e * 2 + nslot) * NB + k]);
                        else if (noise_std > 0.0f) v = __fadd_rn(v, noise_std * z[j]);
                    }
                    sm.scan[k] = __float_as_int(v);
                    if (SS == 1) {
                        orow[k] = v;
                    } else {
                        // _stack_scan (env.py:257-279): [pads = current scan | previous scans,
                        // oldest first | current scan]; the previous observation row still holds
                        // them one slot to the right
                        const int hist = pass == PASS_RESET ? 0 : min(steps, SS - 1);
                        for (int j = 0; j < SS - 1; j++) {
                            if (j < SS - 1 - hist) orow[j * NB + k] = v;
                            else if (pass == PASS_STEP) orow[j * NB + k] = orow[(j + 1) * NB + k];
                        }
                        orow[(SS - 1) * NB + k] = v;
                    }
                    const float thr_k = a.thr[k];
                    c_any |= v < thr_k;
                    d_any |= v < a.dthr[k];
                    margin = fminf(margin, v - thr_k);
                }
            }
        }

        PROF_MARK(5);
        if (pass == PASS_STEP) {
            // ---- reward / done / info on this observation (env.py:464-589)
            int crash, discomf;
            if (WPE > 1) {
                crash = __syncthreads_or(c_any);
                discomf = __syncthreads_or(d_any) && !crash;
            } else {
                crash = __any_sync(FULL, c_any);
                discomf = __any_sync(FULL, d_any) && !crash;
            }
            double mn = CUDART_INF;
            if (discomf) {
                for (int i = 0; i < BPL; i++) {
                    const int k = BEAM(i);
                    const float thr_k = a.thr[k], dthr_k = a.dthr[k];
                    const float den = __fadd_rn(__fsub_rn(dthr_k, thr_k), 1e-6f);
                    mn = fmin(mn, __ddiv_rn(__dsub_rn((double)__int_as_float(sm.scan[k]), (double)thr_k), (double)den));
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(FULL, mn, o));
                if (WPE > 1) {
                    if (lane == 0) sm.red[warp] = mn;
                    __syncthreads();
                }
            }
            if (warp == 0) {
                if (WPE > 1 && discomf)
                    for (int i = 0; i < WPE; i++) mn = fmin(mn, sm.red[i]);
                // lanes 0 / 1: distance to goal from pose / prev_pose
                const double gx = sm.gx, gy = sm.gy;
                const double qx = lane == 0 ? sm.px : sm.ppx, qy = lane == 0 ? sm.py : sm.ppy;
                const double ddx = __dsub_rn(gx, qx), ddy = __dsub_rn(gy, qy);
                const double dq = sqrt(__dadd_rn(__dmul_rn(ddx, ddx), __dmul_rn(ddy, ddy)));
                const double dist_g = __shfl_sync(FULL, dq, 0), pdist = __shfl_sync(FULL, dq, 1);
                __syncwarp();  // lane 1 has read sm.ppx / sm.ppy before lane 0 may rewrite the pose below
                const int success = dist_g < a.dist_thresh;
                const int trunc = a.max_episode_steps > 0 && sm.steps >= a.max_episode_steps && !(success || crash);
                const int done = success || crash || trunc;
                if (lane == 0) {
                    const double pv = sm.pv, pw = sm.pw;
                    double r_s = success ? __dmul_rn(__dmul_rn(1.0, a.r_success), a.r_scale) : 0.0;
                    double r_c = crash ? __dmul_rn(__dmul_rn(-1.0, a.r_crash), a.r_scale) : 0.0;
                    double r_p = __dmul_rn(__dmul_rn(__dsub_rn(pdist, dist_g), a.r_progress), a.r_scale);
                    double r_f = __dmul_rn(__dmul_rn(pv, a.r_forward), a.r_scale);
                    double r_r = __dmul_rn(__dmul_rn(__dmul_rn(-1.0, __dmul_rn(pw, pw)), a.r_rotation), a.r_scale);
                    double r_d = 0.0;
                    if (discomf) r_d = __dmul_rn(__dmul_rn(-__dsub_rn(1.0, mn), a.r_discomfort), a.r_scale);
                    double rew = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(r_s, r_c), r_p), r_f), r_r), r_d);
                    a.reward[e] = (float)rew;
                    a.done[e] = (uint8_t)done;
                    if (a.reward_mirror) a.reward_mirror[e] = (float)rew;
                    if (a.done_mirror) a.done_mirror[e] = (uint8_t)done;
                    a.is_success[e] = (uint8_t)success;
                    a.is_crash[e] = (uint8_t)crash;
                    if (a.truncated) a.truncated[e] = (uint8_t)trunc;
                    a.distance[e] = (float)dist_g;
                    int next = PASS_END;
                    if (done && a.auto_reset) {
                        // auto-reset: draw a spawn tuple (and a map) for the next episode
                        uint4 rnd = philox4x32_10(make_uint4((uint32_t)(a.env_offset + e), (uint32_t)sm.episode, 0x5eedu, 0xfffffff0u),
                                                  make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32)));
                        int nmap = sm.map;
                        if (a.resample_map && a.num_maps > 1) nmap = (int)(((uint64_t)rnd.y * (uint64_t)a.num_maps) >> 32);
                        const navgym_map_t m2 = a.maps[nmap];
                        if (m2.spawn_count > 0) {
                            const long long row = m2.spawn_offset + (long long)(((uint64_t)rnd.x * (uint64_t)m2.spawn_count) >> 32);
                            const double *sp = a.spawn_pool + row * 5;
                            sm.px = sp[0]; sm.py = sp[1]; sm.gx = sp[2]; sm.gy = sp[3]; sm.th = sp[4];
                            sm.map = nmap;
                        } else {  // no pool: restart from the rolled-back pose
                            sm.px = sm.ppx; sm.py = sm.ppy; sm.th = sm.pyaw;
                        }
                        sm.noise_std = a.noise_lo + (a.noise_hi - a.noise_lo) * u01(rnd.z);
                        sm.episode += 1;
                        sm.steps = 0;
                        sm.ppx = sm.px; sm.ppy = sm.py; sm.pv = 0; sm.pw = 0; sm.act_v = 0; sm.act_w = 0;
                        next = PASS_RESET;
                    } else if (crash) {  // env.py:707-717: back to the pose / yaw of prev_obs
                        sm.px = sm.ppx; sm.py = sm.ppy; sm.th = sm.pyaw;
                        next = PASS_RESCAN;
                    }
                    sm.next_pass = next;
                }
            }
            if (WPE > 1) __syncthreads(); else __syncwarp();
            pass = sm.next_pass;
            if (pass == PASS_RESET && a.discs_reset != nullptr) {
                nd = min(a.ndisc_reset[e], a.max_disc);
                ns = a.segs_reset ? min(a.nseg_reset[e], a.max_seg) : 0;
            }
            if (pass != PASS_END) continue;
        }
        break;
    }

    PROF_MARK(6);
    // File this environment under its cost class for the next step: the cycles its last scan
    // took (an auto-reset first scan is taken at the pose the next step starts from), doubled
    // when the next step is likely to end the episode and run a second scan -- the robot is
    // within one step of a crash threshold, of the goal, or of the step limit.  Such
    // environments then start first instead of stretching the end of the launch.  The slot in
    // the class list is claimed here and filled in at the very end, so the atomic's round trip
    // overlaps the epilogue.
    int sched_pos = 0, sched_b = 0;
    if (sched_cnt) {
        bool risky = margin < NAVGYM_RISK_MARGIN;
        if (tid == 0) {
            const double gx = sm.gx - sm.px, gy = sm.gy - sm.py;
            risky |= gx * gx + gy * gy < (a.dist_thresh + NAVGYM_RISK_MARGIN) * (a.dist_thresh + NAVGYM_RISK_MARGIN);
            risky |= a.max_episode_steps > 0 && sm.steps + 1 >= a.max_episode_steps;
        }
        risky = (WPE > 1 ? __syncthreads_or(risky) : __any_sync(FULL, risky)) && a.auto_reset;
        if (tid == 0) {
            const long long kc = ((clock64() - t_pass) << (risky ? 1 : 0)) >> 13;
            sched_b = NAVGYM_SCHED_BUCKETS - 1 - (int)(kc > NAVGYM_SCHED_BUCKETS - 1 ? NAVGYM_SCHED_BUCKETS - 1 : kc);
            sched_pos = atomicAdd(&sched_cnt[sched_b], 1);
        }
    }
    // ---------------- epilogue (warp 0): observation tail + state (env.py:455, 725-727) ---
    if (warp == 0) {
        double yaw;
        if (WPE > 1 && sm.th_spec == sm.th) {
            yaw = sm.yaw_spec;  // warp 1 had the final heading right
        } else {
            double sn, cn;
            sincos(sm.th, &sn, &cn);
            yaw = atan2(sn, cn);  // utils.py:5-9
        }
        double tv = 0.0;
        switch (lane) {
        case 0: tv = sm.ppx; break;
        case 1: tv = sm.ppy; break;
        case 2: tv = sm.px; break;
        case 3: tv = sm.py; break;
        case 4: tv = sm.pv; break;
        case 5: tv = sm.pw; break;
        case 6: tv = yaw; break;
        }
        if (lane < 7) {
            orow[(a.num_scan_stack > 1 ? a.num_scan_stack : 1) * NB + lane] = (float)tv;
            if (a.tail64) a.tail64[(size_t)e * 7 + lane] = tv;
        }
        double nv = 0.0;
        switch (lane) {
        case NAVGYM_S_PX: nv = sm.px; break;
        case NAVGYM_S_PY: nv = sm.py; break;
        case NAVGYM_S_TH: nv = sm.th; break;
        case NAVGYM_S_GX: nv = sm.gx; break;
        case NAVGYM_S_GY: nv = sm.gy; break;
        case NAVGYM_S_PPX: nv = sm.px; break;
        case NAVGYM_S_PPY: nv = sm.py; break;
        case NAVGYM_S_PYAW: nv = yaw; break;
        case NAVGYM_S_PV: nv = sm.act_v; break;
        case NAVGYM_S_PW: nv = sm.act_w; break;
        }
        if (lane < NAVGYM_NS) ST(lane) = nv;
        if (lane == 0) {
            a.steps[e] = sm.steps;
            a.map_id[e] = sm.map;
            if (a.episodes) a.episodes[e] = sm.episode;
            if (a.noise_std) a.noise_std[e] = sm.noise_std;
        }
    }
    if (sched_cnt && tid == 0) sched_list[(size_t)sched_b * B + sched_pos] = e;
    PROF_MARK(7);
#ifdef NAVGYM_PROFILE
    if (tid == 0 && a.tail64) {  // CTA timeline (global ns clock, SM id) for tail analysis
        unsigned long long t_end;
        unsigned smid;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        a.tail64[(size_t)e * 7 + 0] = (double)_g_begin;
        a.tail64[(size_t)e * 7 + 1] = (double)t_end;
        a.tail64[(size_t)e * 7 + 2] = (double)smid;
        a.tail64[(size_t)e * 7 + 3] = (double)blockIdx.x;
    }
#endif
#undef ST
#undef BEAM
}

// ------------------------------------------------------------------ HER batch API
// compute_rewards / compute_terminals / compute_info on a batch of stored observations
// (env.py:464-589: the reference's hindsight-relabelling entry points).  One warp per
// observation row [scan(512) | prev_pose(2) pose(2) vel(2) yaw(1)], goals given separately.
__global__ void __launch_bounds__(256) her_kernel(const navgym_her_args_t a)
{
    const int lane = threadIdx.x & 31;
    const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (n >= a.count) return;
    const unsigned FULL = 0xffffffffu;
    const int SS = a.num_scan_stack > 1 ? a.num_scan_stack : 1;
    const float *o = a.obs + (size_t)n * a.obs_stride + (size_t)(SS - 1) * NB;  // the newest scan
    bool c_any = false, d_any = false;
    double mn = CUDART_INF;
#pragma unroll 4
    for (int i = 0; i < NB / 32; i++) {
        const int k = lane + 32 * i;
        const float v = o[k], thr = a.thr[k], dthr = a.dthr[k];
        c_any |= v < thr;
        d_any |= v < dthr;
        const float den = __fadd_rn(__fsub_rn(dthr, thr), 1e-6f);
        mn = fmin(mn, __ddiv_rn(__dsub_rn((double)v, (double)thr), (double)den));
    }
    const int crash = __any_sync(FULL, c_any);
    const int discomf = __any_sync(FULL, d_any) && !crash;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mn = fmin(mn, __shfl_xor_sync(FULL, mn, off));
    if (lane == 0) {
        const double gx = (double)a.goals[2 * n], gy = (double)a.goals[2 * n + 1];
        const double ppx = (double)o[NB], ppy = (double)o[NB + 1], px = (double)o[NB + 2], py = (double)o[NB + 3];
        const double pv = (double)o[NB + 4], pw = (double)o[NB + 5];
        double dxg = __dsub_rn(gx, px), dyg = __dsub_rn(gy, py);
        double dist = sqrt(__dadd_rn(__dmul_rn(dxg, dxg), __dmul_rn(dyg, dyg)));
        double dxp = __dsub_rn(gx, ppx), dyp = __dsub_rn(gy, ppy);
        double pdist = sqrt(__dadd_rn(__dmul_rn(dxp, dxp), __dmul_rn(dyp, dyp)));
        const int success = dist < a.dist_thresh;
        double r_s = success ? __dmul_rn(__dmul_rn(1.0, a.r_success), a.r_scale) : 0.0;
        double r_c = crash ? __dmul_rn(__dmul_rn(-1.0, a.r_crash), a.r_scale) : 0.0;
        double r_p = __dmul_rn(__dmul_rn(__dsub_rn(pdist, dist), a.r_progress), a.r_scale);
        double r_f = __dmul_rn(__dmul_rn(pv, a.r_forward), a.r_scale);
        double r_r = __dmul_rn(__dmul_rn(__dmul_rn(-1.0, __dmul_rn(pw, pw)), a.r_rotation), a.r_scale);
        double r_d = discomf ? __dmul_rn(__dmul_rn(-__dsub_rn(1.0, mn), a.r_discomfort), a.r_scale) : 0.0;
        if (a.reward)
            a.reward[n] = (float)__dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(r_s, r_c), r_p), r_f), r_r), r_d);
        if (a.done) a.done[n] = (uint8_t)(success || crash);
        if (a.is_success) a.is_success[n] = (uint8_t)success;
        if (a.is_crash) a.is_crash[n] = (uint8_t)crash;
        if (a.distance) a.distance[n] = (float)dist;
    }
}

// ------------------------------------------------------------------ pedestrian lidar
// The scan every simulated pedestrian takes of its surroundings (env.py:683-693): map raycast
// from its own pose + the closed footprints of the robot and the other pedestrians, clipped,
// no noise.  One CTA per agent, threads across beams; the same canonical march / segment
// arithmetic as the robot's scan.
// World-frame closed footprint of an agent at (x, y, th) as 4 segments (env.py:408-414):
// float64 rotation + translation, vertices rounded to float32.
__device__ __forceinline__ void footprint_segments(double x, double y, double th, const double *fp, float4 *out)
{
    const double c = cos(th), s = sin(th);
    float wx[4], wy[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        wx[i] = (float)(c * fp[2 * i] - s * fp[2 * i + 1] + x);
        wy[i] = (float)(s * fp[2 * i] + c * fp[2 * i + 1] + y);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) out[i] = make_float4(wx[i], wy[i], wx[(i + 1) & 3], wy[(i + 1) & 3]);
}

#define NAVGYM_SCAN_SEGS 128  // nearby-footprint segments kept per agent in crowd mode
__global__ void __launch_bounds__(128) agent_scan_kernel(const navgym_scan_args_t a)
{
    __shared__ float4 near_segs[NAVGYM_SCAN_SEGS];
    __shared__ int n_near;
    const int n = blockIdx.x;
    const int e = n / a.agents_per_env, slot = n - e * a.agents_per_env;
    if (a.env_mask && !a.env_mask[e]) return;
    const int live = a.nagent ? min(a.nagent[e], a.agents_per_env) : a.agents_per_env;
    if (slot >= live) return;
    const double *p = a.pose + (size_t)n * 3;
    const float lx = (float)p[0], ly = (float)p[1], lt = (float)p[2];  // env.py:386
    const navgym_map_t m = a.maps[a.map_id[e]];
    const float *dist = a.edt_pool + m.edt_offset;
    const int ci = xy_to_cell(lx, m.ox, m.res, m.H, a.cell_rule);
    const int cj = xy_to_cell(ly, m.oy, m.res, m.W, a.cell_rule);
    const float max_range = (float)((double)m.W * (double)m.H);
    const float t_stop = a.t_stop > 0.0f ? fminf(a.t_stop, max_range) : max_range;
    const float res32 = (float)m.res;
    int ns = 0, s0 = -1, s1 = -1;
    const float4 *segs = nullptr;
    if (a.segs) {
        ns = a.nseg ? min(a.nseg[e], a.max_seg) : 0;
        segs = reinterpret_cast<const float4 *>(a.segs) + (size_t)e * a.max_seg;
        if (a.skip) { s0 = a.skip[2 * (size_t)n]; s1 = s0 + a.skip[2 * (size_t)n + 1]; }
    } else if (a.robot_state) {
        // crowd mode: thread 0 = the robot, thread 1 + j = agent j of this environment
        if (threadIdx.x == 0) n_near = 0;
        __syncthreads();
        const int o = threadIdx.x;
        if (o <= live && o != slot + 1 && 4 * (live + 1) <= NAVGYM_SCAN_SEGS) {
            double ox, oy, oth;
            const double *fp;
            if (o == 0) {
                const size_t B = (size_t)a.num_envs;
                ox = a.robot_state[NAVGYM_S_PX * B + e]; oy = a.robot_state[NAVGYM_S_PY * B + e];
                oth = a.robot_state[NAVGYM_S_TH * B + e];
                fp = a.robot_fp;
            } else {
                const double *q = a.pose + ((size_t)e * a.agents_per_env + (o - 1)) * 3;
                ox = q[0]; oy = q[1]; oth = q[2];
                fp = a.agent_fp;
            }
            float reach = 0.0f;  // farthest footprint vertex from the body origin
#pragma unroll
            for (int i = 0; i < 4; i++) reach = fmaxf(reach, hypotf((float)fp[2 * i], (float)fp[2 * i + 1]));
            const float dc = hypotf((float)ox - lx, (float)oy - ly);
            if (dc <= a.range_max + reach + 0.01f) {
                const int at = atomicAdd(&n_near, 4);
                footprint_segments(ox, oy, oth, fp, near_segs + at);
            }
        }
        __syncthreads();
        ns = n_near;
        segs = near_segs;
    }
    for (int k = threadIdx.x; k < a.num_beams; k += blockDim.x) {
        const float h = (float)__dadd_rn(a.lin[k], (double)lt);
        double sd, cd;
        dir_sincos((double)h, sd, cd);
        const float dx = (float)cd, dy = (float)sd;
        int hx, hy;
        float r = __fmul_rn(march(dist, m.W, m.H, (float)ci, (float)cj, dx, dy, max_range, t_stop, hx, hy), res32);
        for (int s = 0; s < ns; s++) {
            if (s >= s0 && s < s1) continue;
            const float4 sg = segs[s];
            r = fminf(r, seg_hit(lx, ly, dx, dy, sg.x, sg.y, sg.z, sg.w));
        }
        a.ranges[(size_t)n * a.num_beams + k] = fminf(fmaxf(r, 0.0f), a.range_max);
    }
}

// ------------------------------------------------------------------ pedestrian routes
// (include/navgym_b200.h, navgym_plan_args_t.)  One thread per pedestrian.
__device__ __forceinline__ int plan_cell(double v, double origin, double res, int dim)
{
    int c = (int)floor((v - origin) / res);
    return c < 0 ? 0 : (c >= dim ? dim - 1 : c);
}

// From cell (cx, cy) walk the field downhill until the cell centre is more than 2 m from the
// starting point (sx, sy) or the goal cell is reached.  Returns false on an unreachable cell.
__device__ __forceinline__ bool plan_advance(const uint16_t *f, const navgym_plan_map_t &m, int cx, int cy,
                                             double sx, double sy, double &wx, double &wy, bool &is_goal)
{
    unsigned d = f[cy * m.W + cx];
    is_goal = false;
    if (d == 65535u) return false;
    for (int it = 0; it < 64; it++) {
        if (d == 0u) { is_goal = true; break; }
        int bx = cx, by = cy;
        unsigned bd = d;
        if (cx > 0 && f[cy * m.W + cx - 1] < bd) { bd = f[cy * m.W + cx - 1]; bx = cx - 1; by = cy; }
        if (cx + 1 < m.W && f[cy * m.W + cx + 1] < bd) { bd = f[cy * m.W + cx + 1]; bx = cx + 1; by = cy; }
        if (cy > 0 && f[(cy - 1) * m.W + cx] < bd) { bd = f[(cy - 1) * m.W + cx]; bx = cx; by = cy - 1; }
        if (cy + 1 < m.H && f[(cy + 1) * m.W + cx] < bd) { bd = f[(cy + 1) * m.W + cx]; bx = cx; by = cy + 1; }
        if (bd >= d) break;  // local minimum that is not the goal: cannot happen on a BFS field
        cx = bx; cy = by; d = bd;
        wx = m.ox + (cx + 0.5) * m.res;
        wy = m.oy + (cy + 0.5) * m.res;
        if ((wx - sx) * (wx - sx) + (wy - sy) * (wy - sy) > 4.0) return true;
    }
    wx = m.ox + (cx + 0.5) * m.res;
    wy = m.oy + (cy + 0.5) * m.res;
    return true;
}

__global__ void peds_plan_kernel(const navgym_plan_args_t a)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= a.num_envs * a.max_ped) return;
    const int e = n / a.max_ped, slot = n - e * a.max_ped;
    if (a.nped && slot >= a.nped[e]) return;
    const navgym_plan_map_t m = a.maps[a.map_id[e]];
    if (a.respawn && a.respawn[e] && a.cand_pose) {
        // env.py:785-806: the pedestrian drawn for this episode by the previous call
        for (int i = 0; i < 3; i++) a.pose_rw[3 * (size_t)n + i] = a.cand_pose[3 * (size_t)n + i];
        a.v_pref[n] = a.cand_v_pref[n];
        a.has_legs[n] = a.cand_legs[n];
        a.goal_id[n] = a.cand_goal[n];
        a.waypoint[2 * (size_t)n] = CUDART_NAN;
        a.waypoint[2 * (size_t)n + 1] = CUDART_NAN;
        for (int i = 0; i < 3; i++) a.dist_travelled[3 * (size_t)n + i] = 0.0;
        a.vel[2 * (size_t)n] = a.vel[2 * (size_t)n + 1] = 0.0;
        a.prev_action[2 * (size_t)n] = a.prev_action[2 * (size_t)n + 1] = 0.0f;
    }
    const double px = a.pose[3 * (size_t)n], py = a.pose[3 * (size_t)n + 1], th = a.pose[3 * (size_t)n + 2];
    const int cx = plan_cell(px, m.ox, m.res, m.W), cy = plan_cell(py, m.oy, m.res, m.H);
    const size_t fsz = (size_t)m.W * m.H;
    int g = a.goal_id[n];
    const double *goals = a.goals + 2 * m.goal_offset;
    double gx = goals[2 * g], gy = goals[2 * g + 1];
    double wx = a.waypoint[2 * (size_t)n], wy = a.waypoint[2 * (size_t)n + 1];
    // 1. arrived (env.py:666-668): a new goal, if one qualifies
    if ((px - gx) * (px - gx) + (py - gy) * (py - gy) < 0.25) {
        const uint4 r = philox4x32_10(make_uint4((uint32_t)(a.env_offset + e), (uint32_t)slot, (uint32_t)a.step, 0x9ed5u),
                                      make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32)));
        const uint32_t draws[4] = {r.x, r.y, r.z, r.w};
        for (int i = 0; i < 4; i++) {
            const int c = (int)(((uint64_t)draws[i] * (uint64_t)m.num_goals) >> 32);
            const double qx = goals[2 * c], qy = goals[2 * c + 1];
            const double dd = (qx - px) * (qx - px) + (qy - py) * (qy - py);
            if (c != g && dd > a.min_goal_dist * a.min_goal_dist &&
                a.fields[m.field_offset + c * fsz + (size_t)cy * m.W + cx] != 65535u) {
                g = c; gx = qx; gy = qy;
                wx = CUDART_NAN;
                break;
            }
        }
        a.goal_id[n] = g;
    }
    const uint16_t *f = a.fields + m.field_offset + g * fsz;
    // 2. / 3. the waypoint: first one from here, later ones from the previous waypoint
    bool is_goal = (wx == gx) & (wy == gy);
    if (wx != wx) {
        if (!plan_advance(f, m, cx, cy, px, py, wx, wy, is_goal)) { wx = gx; wy = gy; is_goal = true; }
        if (is_goal) { wx = gx; wy = gy; }
    }
    for (int it = 0; it < 8 && !is_goal && (px - wx) * (px - wx) + (py - wy) * (py - wy) < 1.0; it++) {
        const double sx = wx, sy = wy;
        if (!plan_advance(f, m, plan_cell(sx, m.ox, m.res, m.W), plan_cell(sy, m.oy, m.res, m.H), sx, sy, wx, wy, is_goal)) {
            is_goal = true;
        }
        if (is_goal) { wx = gx; wy = gy; }
    }
    a.waypoint[2 * (size_t)n] = wx;
    a.waypoint[2 * (size_t)n + 1] = wy;
    // env.py:641-645: the goal in the pedestrian's frame
    const double c = cos(th), s = sin(th);
    a.goal_local[2 * (size_t)n] = (float)((wx - px) * c + (wy - py) * s);
    a.goal_local[2 * (size_t)n + 1] = (float)(-(wx - px) * s + (wy - py) * c);

    // ---- the pedestrian of this slot in the environment's next episode
    if (a.cand_pose) {
        const size_t B = (size_t)a.num_envs;
        const uint32_t ge = (uint32_t)(a.env_offset + e);
        double rx = a.robot_state[NAVGYM_S_PX * B + e], ry = a.robot_state[NAVGYM_S_PY * B + e];
        int nmap = a.map_id[e];
        if (a.cand_next_spawn && a.robot_maps) {
            // the spawn tuple step_kernel draws on auto-reset (keep in step with it)
            const uint4 rnd = philox4x32_10(make_uint4(ge, (uint32_t)(a.episodes ? a.episodes[e] : 0), 0x5eedu, 0xfffffff0u),
                                            make_uint2((uint32_t)a.robot_seed, (uint32_t)(a.robot_seed >> 32)));
            int cand_map = nmap;
            if (a.resample_map && a.num_maps > 1) cand_map = (int)(((uint64_t)rnd.y * (uint64_t)a.num_maps) >> 32);
            const navgym_map_t m2 = a.robot_maps[cand_map];
            if (m2.spawn_count > 0) {
                const long long row = m2.spawn_offset + (long long)(((uint64_t)rnd.x * (uint64_t)m2.spawn_count) >> 32);
                rx = a.spawn_pool[row * 5];
                ry = a.spawn_pool[row * 5 + 1];
                nmap = cand_map;
            }
        }
        const navgym_plan_map_t mc = a.maps[nmap];
        const uint2 key = make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32));
        const uint4 r0 = philox4x32_10(make_uint4(ge, (uint32_t)slot, (uint32_t)a.step, 0x5b0au), key);
        const uint4 r1 = philox4x32_10(make_uint4(ge, (uint32_t)slot, (uint32_t)a.step, 0x5b0bu), key);
        const uint4 r2 = philox4x32_10(make_uint4(ge, (uint32_t)slot, (uint32_t)a.step, 0x5b0cu), key);
        const uint32_t draws[6] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y};
        double sx = rx, sy = ry;
        for (int i = 0; i < 6 && mc.free_count > 0; i++) {
            const long long row = mc.free_offset + (long long)(((uint64_t)draws[i] * (uint64_t)mc.free_count) >> 32);
            sx = a.free_xy[2 * row]; sy = a.free_xy[2 * row + 1];
            if ((sx - rx) * (sx - rx) + (sy - ry) * (sy - ry) >= a.min_robot_dist * a.min_robot_dist) break;
        }
        const double cth = 6.283185307179586 * (double)u01(r2.x);
        const bool legs = (double)u01(r2.z) < a.has_legs_ratio;
        a.cand_pose[3 * (size_t)n] = sx;
        a.cand_pose[3 * (size_t)n + 1] = sy;
        a.cand_pose[3 * (size_t)n + 2] = cth;
        a.cand_v_pref[n] = a.v_pref_lo + (a.v_pref_hi - a.v_pref_lo) * (double)u01(r2.y);
        a.cand_legs[n] = legs;
        a.cand_goal[n] = (int)(((uint64_t)r2.w * (uint64_t)mc.num_goals) >> 32);
        float *q = a.cand_rows + (size_t)n * NAVGYM_PED_F;
        q[0] = (float)sx; q[1] = (float)sy; q[2] = (float)cth;
        q[9] = 0.0f; q[10] = 0.0f; q[11] = 0.0f;
        q[12] = legs ? 1.0f : 0.0f;
    }
}

// ------------------------------------------------------------------ pedestrian motion
// (include/navgym_b200.h, navgym_move_args_t.)  One thread per pedestrian.
__global__ void peds_move_kernel(const navgym_move_args_t a)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= a.num_envs * a.max_ped) return;
    const int e = n / a.max_ped, slot = n - e * a.max_ped;
    if (a.nped && slot >= a.nped[e]) return;
    // env.py:655-661
    const float m0 = fminf(fmaxf(a.mean[2 * (size_t)n], 0.0f), 1.0f);
    const float m1 = fminf(fmaxf(a.mean[2 * (size_t)n + 1], -1.0f), 1.0f);
    a.prev_action[2 * (size_t)n] = m0;
    a.prev_action[2 * (size_t)n + 1] = m1;
    const double factor = a.v_pref[n];
    const double v = (double)m0 * factor, w = (double)m1 * factor;
    // human.py:32-41
    double x = a.pose[3 * (size_t)n], y = a.pose[3 * (size_t)n + 1];
    const double th0 = a.pose[3 * (size_t)n + 2];
    const double vx = v * cos(th0), vy = v * sin(th0);
    const double th1 = th0 + w * a.dt;
    x = x + cos(th1) * v * a.dt;
    y = y + sin(th1) * v * a.dt;
    const double twopi = 6.283185307179586;
    double th = fmod(th1, twopi);
    if (th != 0 && th < 0) th += twopi;
    a.pose[3 * (size_t)n] = x;
    a.pose[3 * (size_t)n + 1] = y;
    a.pose[3 * (size_t)n + 2] = th;
    a.vel[2 * (size_t)n] = vx;
    a.vel[2 * (size_t)n + 1] = vy;
    // env.py:237-255: rotation rate from the previous observation's yaw, world velocity into
    // the base frame (pose2d inverse_pose2d / apply_tf_to_vel written out), integrated
    const double prev_yaw = atan2(sin(th0), cos(th0));
    const double vrot = (th - prev_yaw) / a.dt;
    const double c = cos(th), s = sin(th);
    double *d = a.dist_travelled + 3 * (size_t)n;
    d[0] += (c * vx + s * vy) * a.dt;
    d[1] += (-s * vx + c * vy) * a.dt;
    d[2] += vrot * a.dt;
    float *q = a.rows + (size_t)n * NAVGYM_PED_F;
    q[0] = (float)x; q[1] = (float)y; q[2] = (float)th;
    q[9] = (float)d[0]; q[10] = (float)d[1]; q[11] = (float)d[2];
    q[12] = a.has_legs[n] ? 1.0f : 0.0f;
}

// ------------------------------------------------------------------ pedestrian policy front end
// (include/navgym_b200.h, navgym_policy_features.)  128 threads; a CTA keeps the second
// convolution's weights in shared memory and walks over pedestrians: conv1 fills h1 in shared
// memory, then conv2 is register-tiled: warp w owns output channels 8w .. 8w + 7, lane l output
// positions 4l .. 4l + 3 (32 accumulators); per input channel a lane reads its 9 inputs with
// three loads and the warp's 24 weights as six broadcast LDS.128 -- 9 shared-memory loads per
// 96 FMA, where one position per thread needed 25.
#define PF_H1 264  // row pitch of h1 (32-byte multiple): [0] = left pad, [1 + q] = conv1 output q (q < 255), [256] = right pad
__global__ void __launch_bounds__(128) policy_features_kernel(const float *__restrict__ scan, int n,
                                                              const float *__restrict__ w1, const float *__restrict__ b1,
                                                              const float *__restrict__ w2, const float *__restrict__ b2,
                                                              float *__restrict__ out)
{
    __shared__ __align__(16) float w2s[32 * 3 * 32];  // [ci][tap][co]
    __shared__ __align__(16) float h1[32 * PF_H1];
    __shared__ float xs[516];                          // [0] = left pad, [1 + i] = input i
    __shared__ float w1s[32 * 5], b1s[32], b2s[32];
    const int t = threadIdx.x;
    for (int i = t; i < 32 * 32 * 3; i += 128) {
        const int co = i / 96, ci = (i / 3) % 32, k = i % 3;
        w2s[(ci * 3 + k) * 32 + co] = w2[i];
    }
    for (int i = t; i < 160; i += 128) w1s[i] = w1[i];
    if (t < 32) { b1s[t] = b1[t]; b2s[t] = b2[t]; }
    for (int c = t; c < 32; c += 128) { h1[c * PF_H1] = 0.0f; h1[c * PF_H1 + 256] = 0.0f; }
    if (t == 0) { xs[0] = 0.0f; xs[513] = 0.0f; xs[514] = 0.0f; xs[515] = 0.0f; }
    for (int ped = blockIdx.x; ped < n; ped += gridDim.x) {
        __syncthreads();  // weights ready / previous pedestrian's h1 no longer read
        for (int i = t; i < 512; i += 128) {
            const double r = fmin(fmax((double)scan[(size_t)ped * 512 + i], 0.0), 6.0);
            xs[1 + i] = (float)(r / 6.0 - 0.5);  // env.py:627-629, 648
        }
        __syncthreads();
        {   // conv1: thread t owns output positions t and t + 128 of every channel
            float xa[5], xb[5];
#pragma unroll
            for (int k = 0; k < 5; k++) { xa[k] = xs[2 * t + k]; xb[k] = t < 127 ? xs[2 * (t + 128) + k] : 0.0f; }  // input 2q - 1 + k
#pragma unroll 4
            for (int c = 0; c < 32; c++) {
                float a0 = b1s[c], a1 = a0;
#pragma unroll
                for (int k = 0; k < 5; k++) { a0 = fmaf(w1s[c * 5 + k], xa[k], a0); a1 = fmaf(w1s[c * 5 + k], xb[k], a1); }
                h1[c * PF_H1 + 1 + t] = fmaxf(a0, 0.0f);
                if (t < 127) h1[c * PF_H1 + 129 + t] = fmaxf(a1, 0.0f);
            }
        }
        __syncthreads();
        const int wco = (t >> 5) * 8, p0 = (t & 31) * 4;
        float acc[8][4];
#pragma unroll
        for (int c = 0; c < 8; c++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[c][j] = b2s[wco + c];
#pragma unroll 2
        for (int ci = 0; ci < 32; ci++) {
            // inputs 2 p0 - 1 .. 2 p0 + 7 = h1 row entries 2 p0 .. 2 p0 + 8
            const float *row = h1 + ci * PF_H1 + 2 * p0;
            const float4 i0 = *reinterpret_cast<const float4 *>(row), i1 = *reinterpret_cast<const float4 *>(row + 4);
            const float in[9] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w, row[8]};
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const float4 *wr = reinterpret_cast<const float4 *>(w2s + (ci * 3 + k) * 32 + wco);
                const float4 wa = wr[0], wb = wr[1];
                const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
                for (int c = 0; c < 8; c++)
#pragma unroll
                    for (int j = 0; j < 4; j++) acc[c][j] = fmaf(in[2 * j + k], wv[c], acc[c][j]);
            }
        }
        float *o = out + (size_t)ped * 4096 + p0;
#pragma unroll
        for (int c = 0; c < 8; c++)
            *reinterpret_cast<float4 *>(o + (wco + c) * 128) =
                make_float4(fmaxf(acc[c][0], 0.0f), fmaxf(acc[c][1], 0.0f), fmaxf(acc[c][2], 0.0f), fmaxf(acc[c][3], 0.0f));
    }
}

// ------------------------------------------------------------------ scripted pedestrians
// Pedestrian motion + geometry for the batched simulator (scripted stand-in
// for the reference's CNN-driven humans whose weights are absent): each pedestrian walks
// between two waypoints at its preferred speed with Human.set_vel's unicycle update
// (human.py:32-41), turning at <= 1 rad/s toward the current target, and its leg-gait odometry
// advances as in _update_dist_travelled (env.py:237-255).  Emits what the robot's lidar sees:
// two leg discs (pymap2d CSimAgent "legs") for legged pedestrians, the 0.44 x 0.38 m box
// footprint (human.py:5-10) as four segments otherwise (env.py:398-414), or one trunk disc
// per pedestrian in trunk mode.  One thread per environment.
__global__ void peds_advance_kernel(const navgym_peds_args_t a)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= a.num_envs) return;
    const int P = a.nped ? min(a.nped[e], a.max_ped) : a.max_ped;
    float *pp = a.peds + (size_t)e * a.max_ped * NAVGYM_PED_F;
    float *discs = a.discs + (size_t)e * a.max_disc * 3;
    float *segs = a.segs ? a.segs + (size_t)e * a.max_seg * 4 : nullptr;
    int nd = 0, ns = 0;
    for (int p = 0; p < P; p++) {
        float *q = pp + p * NAVGYM_PED_F;
        float x = q[0], y = q[1], th = q[2];
        const float v = q[3];
        float tgt = q[8];
        if (a.advance) {
            float gx = tgt > 0.5f ? q[6] : q[4], gy = tgt > 0.5f ? q[7] : q[5];
            if ((gx - x) * (gx - x) + (gy - y) * (gy - y) < 0.25f) {  // reached: turn back
                tgt = 1.0f - tgt;
                gx = tgt > 0.5f ? q[6] : q[4];
                gy = tgt > 0.5f ? q[7] : q[5];
            }
            float err = atan2f(gy - y, gx - x) - th;
            err -= 6.2831853f * rintf(err * 0.15915494f);
            const float w = fminf(fmaxf(err / a.dt, -1.0f), 1.0f);
            const float vx = v * cosf(th), vy = v * sinf(th);  // human.py:35-36 (old heading)
            const float thn = th + w * a.dt;
            x += cosf(thn) * v * a.dt;
            y += sinf(thn) * v * a.dt;
            // leg gait odometry in the base frame (env.py:251-255)
            const float c = cosf(thn), s_ = sinf(thn);
            q[9] += (c * vx + s_ * vy) * a.dt;
            q[10] += (-s_ * vx + c * vy) * a.dt;
            q[11] += w * a.dt;
            th = thn - 6.2831853f * floorf(thn * 0.15915494f);
            q[0] = x; q[1] = y; q[2] = th; q[8] = tgt;
        }
        const float c = cosf(th), s_ = sinf(th);
        if (a.trunk_mode) {
            if (nd < a.max_disc) { discs[3 * nd] = x; discs[3 * nd + 1] = y; discs[3 * nd + 2] = q[13]; nd++; }
        } else if (q[12] > 0.5f) {  // legs
            const float front = 0.3f * cosf(q[9] * (2.0f / 0.3f) + q[11]);
            const float side = 0.1f * cosf(q[10] * (2.0f / 0.1f) + q[11]) + 0.1f;
            if (nd + 1 < a.max_disc) {
                discs[3 * nd] = x + c * front - s_ * side; discs[3 * nd + 1] = y + s_ * front + c * side;
                discs[3 * nd + 2] = 0.03f; nd++;
                discs[3 * nd] = x - c * front + s_ * side; discs[3 * nd + 1] = y - s_ * front - c * side;
                discs[3 * nd + 2] = 0.03f; nd++;
            }
        } else if (segs && ns + 3 < a.max_seg) {  // box footprint, closed
            const float fx[4] = {0.22f, -0.22f, -0.22f, 0.22f}, fy[4] = {0.19f, 0.19f, -0.19f, -0.19f};
            float wx[4], wy[4];
#pragma unroll
            for (int i = 0; i < 4; i++) { wx[i] = c * fx[i] - s_ * fy[i] + x; wy[i] = s_ * fx[i] + c * fy[i] + y; }
#pragma unroll
            for (int i = 0; i < 4; i++) {
                float *sg = segs + 4 * (ns + i);
                sg[0] = wx[i]; sg[1] = wy[i]; sg[2] = wx[(i + 1) & 3]; sg[3] = wy[(i + 1) & 3];
            }
            ns += 4;
        }
    }
    a.ndisc[e] = nd;
    if (a.nseg) a.nseg[e] = ns;
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
    dir_sincos((double)h, s, c);
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
    dir_sincos((double)headings[k], s, c);
    float dx = (float)c, dy = (float)s, r = ranges[k];
    for (int i = 0; i < S; i++)
        r = fminf(r, seg_hit(ox, oy, dx, dy, segs[4 * i], segs[4 * i + 1], segs[4 * i + 2], segs[4 * i + 3]));
    for (int i = 0; i < D; i++)
        r = fminf(r, disc_hit(ox, oy, dx, dy, discs[3 * i], discs[3 * i + 1], discs[3 * i + 2]));
    ranges[k] = r;
}

// ------------------------------------------------------------------ C ABI
// Launch shape of the fused kernel: warps per environment (WPE) and march slots per lane.
// NAVGYM_WPE / NAVGYM_SLOTS override the defaults for tuning runs.
static int env_int(const char *name, int dflt)
{
    const char *s = getenv(name);
    return s ? atoi(s) : dflt;
}

template <bool RESET, int WPE, int SLOTS>
static void launch_one(const navgym_step_args_t &a, cudaStream_t st)
{
    const int count = a.env_count > 0 ? a.env_count : a.num_envs - a.env_begin;
    step_kernel<RESET, WPE, SLOTS><<<count, WPE * 32, 0, st>>>(a);
}

template <bool RESET>
static void launch_step(const navgym_step_args_t &a, cudaStream_t st)
{
    static const int wpe = env_int("NAVGYM_WPE", 2), slots = env_int("NAVGYM_SLOTS", 1);
    switch (wpe * 10 + slots) {
    case 11: launch_one<RESET, 1, 1>(a, st); break;
    case 12: launch_one<RESET, 1, 2>(a, st); break;
    case 14: launch_one<RESET, 1, 4>(a, st); break;
    case 22: launch_one<RESET, 2, 2>(a, st); break;
    case 24: launch_one<RESET, 2, 4>(a, st); break;
    case 41: launch_one<RESET, 4, 1>(a, st); break;
    case 44: launch_one<RESET, 4, 4>(a, st); break;
    case 81: launch_one<RESET, 8, 1>(a, st); break;
    case 82: launch_one<RESET, 8, 2>(a, st); break;
    case 42: launch_one<RESET, 4, 2>(a, st); break;
    default: launch_one<RESET, 2, 1>(a, st); break;
    }
    g_launches++;
}

#define CK(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) return (int)_e; } while (0)

extern "C" {

#ifdef NAVGYM_PROFILE
int navgym_debug_read_prof(unsigned long long *out, int reset)
{
    cudaError_t e = cudaMemcpyFromSymbol(out, g_prof, sizeof(unsigned long long) * 16);
    if (!e && reset) { unsigned long long z[16] = {0}; e = cudaMemcpyToSymbol(g_prof, z, sizeof(z)); }
    return (int)e;
}
#endif

int navgym_abi_version(void) { return 1; }
int navgym_sizeof_step_args(void) { return (int)sizeof(navgym_step_args_t); }
int navgym_sizeof_map(void) { return (int)sizeof(navgym_map_t); }
int navgym_sizeof_her_args(void) { return (int)sizeof(navgym_her_args_t); }
int navgym_sizeof_peds_args(void) { return (int)sizeof(navgym_peds_args_t); }
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
    if (args->obs_stride < (args->num_scan_stack > 1 ? args->num_scan_stack : 1) * NB + NAVGYM_OBS_TAIL || args->env_begin < 0 || args->env_begin + args->env_count > args->num_envs)
        return (int)cudaErrorInvalidValue;
    launch_step<false>(*args, (cudaStream_t)stream);
    return (int)cudaGetLastError();
}

// ---- host-buffer step: chunked launches on prioritised streams, D2H of early chunks
// overlapping the raycast of later ones -----------------------------------------------------
#define NAVGYM_MAX_CHUNKS 8
// One submit = H2D(actions) -> step -> 3 x D2H on the group's stream.  Issued call by call that
// is five driver calls per group and step; the sequence only depends on the argument block, the
// host pointers and the schedule phase, so it is captured once per (group, phase) into a CUDA
// graph and replayed with one cudaGraphLaunch while those stay the same.
struct navgym_group_graph {
    cudaGraphExec_t exec;
    int kernels;            // kernel nodes in the graph
    navgym_step_args_t key;
    const void *host[4];
};
struct navgym_host_pipe {
    int chunks, num_envs;
    cudaStream_t streams[NAVGYM_MAX_CHUNKS];
    cudaEvent_t ready;
    int32_t *sched[NAVGYM_MAX_CHUNKS];
    int phase[NAVGYM_MAX_CHUNKS];
    int b0[NAVGYM_MAX_CHUNKS + 1];
    navgym_group_graph graphs[NAVGYM_MAX_CHUNKS][3];
    int use_graphs;
};

navgym_host_pipe_t *navgym_host_pipe_create(int chunks, int num_envs, int longest_first)
{
    if (chunks < 1 || chunks > NAVGYM_MAX_CHUNKS || num_envs < 1) return nullptr;
    navgym_host_pipe_t *p = new navgym_host_pipe_t();
    p->chunks = chunks;
    p->num_envs = num_envs;
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);  // hi is the numerically lowest = highest priority
    for (int c = 0; c <= chunks; c++) p->b0[c] = (int)((long long)num_envs * c / chunks);
    for (int c = 0; c < chunks; c++) {
        static const int spread = env_int("NAVGYM_PIPE_PRIO", 1);
        int prio = spread ? hi + c : lo;
        if (prio > lo) prio = lo;
        if (cudaStreamCreateWithPriority(&p->streams[c], cudaStreamNonBlocking, prio) != cudaSuccess) { delete p; return nullptr; }
        p->sched[c] = nullptr;
        p->phase[c] = 0;
        if (longest_first) {
            const size_t n = 3 * NAVGYM_SCHED_BUCKETS + (size_t)3 * NAVGYM_SCHED_BUCKETS * num_envs;
            int32_t *h = new int32_t[n]();
            const int cnt = p->b0[c + 1] - p->b0[c];
            h[0] = cnt;
            for (int i = 0; i < cnt; i++) h[3 * NAVGYM_SCHED_BUCKETS + i] = p->b0[c] + i;
            cudaError_t err = cudaMalloc(&p->sched[c], n * sizeof(int32_t));
            if (!err) err = cudaMemcpy(p->sched[c], h, n * sizeof(int32_t), cudaMemcpyHostToDevice);
            delete[] h;
            if (err) { delete p; return nullptr; }
        }
    }
    cudaEventCreateWithFlags(&p->ready, cudaEventDisableTiming);
    memset(p->graphs, 0, sizeof(p->graphs));
    p->use_graphs = env_int("NAVGYM_HOST_GRAPHS", 1);
    return p;
}

void navgym_host_pipe_destroy(navgym_host_pipe_t *p)
{
    if (!p) return;
    for (int c = 0; c < p->chunks; c++) {
        cudaStreamSynchronize(p->streams[c]);
        for (int i = 0; i < 3; i++)
            if (p->graphs[c][i].exec) cudaGraphExecDestroy(p->graphs[c][i].exec);
        cudaStreamDestroy(p->streams[c]);
        if (p->sched[c]) cudaFree(p->sched[c]);
    }
    cudaEventDestroy(p->ready);
    delete p;
}

// Device-visible alias of a pinned (mapped) host address, or NULL for pageable memory.
static void *mapped_alias(const void *ptr)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
}
static bool is_pinned_host(const void *ptr) { return mapped_alias(ptr) != nullptr; }

// H2D(actions) -> step -> D2H for environments [a.env_begin, +a.env_count) on `st`.  reward and
// done are a few bytes per environment: when their host arrays are mapped the kernel stores them
// there itself (reward_mirror / done_mirror) and the observation rows are the only D2H copy.
static int enqueue_group(navgym_step_args_t a, cudaStream_t st, const float *actions_host,
                         float *obs_host, float *reward_host, uint8_t *done_host, bool h2d)
{
    const size_t b0 = (size_t)a.env_begin, n = (size_t)a.env_count;
    static const int mirrors = env_int("NAVGYM_HOST_MIRRORS", 1);
    a.reward_mirror = mirrors ? (float *)mapped_alias(reward_host) : nullptr;
    a.done_mirror = mirrors ? (uint8_t *)mapped_alias(done_host) : nullptr;
    if (h2d)
        CK(cudaMemcpyAsync((void *)(a.actions + 2 * b0), actions_host + 2 * b0, n * 2 * sizeof(float),
                           cudaMemcpyHostToDevice, st));
    int err = navgym_step_batch(&a, st);
    if (err) return err;
    CK(cudaMemcpyAsync(obs_host + b0 * a.obs_stride, a.obs + b0 * a.obs_stride,
                       n * a.obs_stride * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (!a.reward_mirror)
        CK(cudaMemcpyAsync(reward_host + b0, a.reward + b0, n * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (!a.done_mirror)
        CK(cudaMemcpyAsync(done_host + b0, a.done + b0, n, cudaMemcpyDeviceToHost, st));
    return 0;
}

int navgym_step_batch_host(navgym_host_pipe_t *p, const navgym_step_args_t *args, void *stream,
                           const float *actions_host, float *obs_host, float *reward_host,
                           uint8_t *done_host)
{
    if (!p || args->num_envs != p->num_envs || !args->actions) return (int)cudaErrorInvalidValue;
    cudaStream_t in = (cudaStream_t)stream;
    const size_t B = (size_t)args->num_envs;
    CK(cudaMemcpyAsync((void *)args->actions, actions_host, B * 2 * sizeof(float), cudaMemcpyHostToDevice, in));
    CK(cudaEventRecord(p->ready, in));
    for (int c = 0; c < p->chunks; c++) {
        cudaStream_t st = p->streams[c];
        navgym_step_args_t a = *args;
        a.env_begin = p->b0[c];
        a.env_count = p->b0[c + 1] - p->b0[c];
        if (a.env_count <= 0) continue;
        a.sched = p->sched[c];
        a.sched_phase = p->phase[c];
        CK(cudaStreamWaitEvent(st, p->ready, 0));
        int err = enqueue_group(a, st, actions_host, obs_host, reward_host, done_host, false);
        if (err) return err;
        if (p->sched[c]) p->phase[c] = (p->phase[c] + 1) % 3;
    }
    for (int c = 0; c < p->chunks; c++) CK(cudaStreamSynchronize(p->streams[c]));
    return 0;
}

// Asynchronous variant for callers that keep several groups of environments in flight
// (group g = the pipe's g-th env range): submit enqueues H2D(actions) -> step -> D2H(results)
// for one group on that group's stream and returns at once; wait blocks until that group's
// results have landed.  While the host consumes group A's observations, group B is stepping.
int navgym_step_batch_host_submit(navgym_host_pipe_t *p, const navgym_step_args_t *args, int group,
                                  const float *actions_host, float *obs_host, float *reward_host,
                                  uint8_t *done_host)
{
    if (!p || group < 0 || group >= p->chunks || args->num_envs != p->num_envs || !args->actions)
        return (int)cudaErrorInvalidValue;
    cudaStream_t st = p->streams[group];
    navgym_step_args_t a = *args;
    a.env_begin = p->b0[group];
    a.env_count = p->b0[group + 1] - p->b0[group];
    if (a.env_count <= 0) return 0;
    a.sched = p->sched[group];
    a.sched_phase = p->phase[group];
    if (p->sched[group]) p->phase[group] = (p->phase[group] + 1) % 3;
    if (!p->use_graphs) return enqueue_group(a, st, actions_host, obs_host, reward_host, done_host, true);

    navgym_group_graph &g = p->graphs[group][a.sched_phase];
    const void *host[4] = {actions_host, obs_host, reward_host, done_host};
    if (!g.exec || memcmp(&g.key, &a, sizeof(a)) != 0 || memcmp(g.host, host, sizeof(host)) != 0) {
        if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
        // pageable host memory cannot be captured: such callers keep the call-by-call path
        for (int i = 0; i < 4; i++)
            if (!is_pinned_host(host[i])) return enqueue_group(a, st, actions_host, obs_host, reward_host, done_host, true);
        cudaGraph_t graph = nullptr;
        CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        const uint64_t launches = g_launches;
        int err = enqueue_group(a, st, actions_host, obs_host, reward_host, done_host, true);
        g.kernels = (int)(g_launches - launches);
        g_launches = launches;  // captured, not launched
        cudaError_t cap = cudaStreamEndCapture(st, &graph);
        if (err || cap != cudaSuccess) {
            if (graph) cudaGraphDestroy(graph);
            return err ? err : (int)cap;
        }
        cap = cudaGraphInstantiate(&g.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (cap != cudaSuccess) { g.exec = nullptr; return (int)cap; }
        g.key = a;
        memcpy(g.host, host, sizeof(host));
    }
    CK(cudaGraphLaunch(g.exec, st));
    g_launches += g.kernels;
    return 0;
}

int navgym_step_batch_host_wait(navgym_host_pipe_t *p, int group)
{
    if (!p || group < 0 || group >= p->chunks) return (int)cudaErrorInvalidValue;
    return (int)cudaStreamSynchronize(p->streams[group]);
}

int navgym_reset_obs_batch(const navgym_step_args_t *args, void *stream)
{
    if (args->num_envs <= 0) return 0;
    if (args->obs_stride < (args->num_scan_stack > 1 ? args->num_scan_stack : 1) * NB + NAVGYM_OBS_TAIL)
        return (int)cudaErrorInvalidValue;
    launch_step<true>(*args, (cudaStream_t)stream);
    return (int)cudaGetLastError();
}

int navgym_compute_rewards(const navgym_her_args_t *args, void *stream)
{
    if (args->count <= 0) return 0;
    if (args->obs_stride < (args->num_scan_stack > 1 ? args->num_scan_stack : 1) * NB + NAVGYM_OBS_TAIL)
        return (int)cudaErrorInvalidValue;
    her_kernel<<<(args->count + 7) / 8, 256, 0, (cudaStream_t)stream>>>(*args);
    g_launches++;
    return (int)cudaGetLastError();
}

int navgym_peds_advance(const navgym_peds_args_t *args, void *stream)
{
    if (args->num_envs <= 0) return 0;
    peds_advance_kernel<<<(args->num_envs + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*args);
    g_launches++;
    return (int)cudaGetLastError();
}

int navgym_sizeof_scan_args(void) { return (int)sizeof(navgym_scan_args_t); }
int navgym_sizeof_plan_args(void) { return (int)sizeof(navgym_plan_args_t); }
int navgym_sizeof_plan_map(void) { return (int)sizeof(navgym_plan_map_t); }
int navgym_sizeof_move_args(void) { return (int)sizeof(navgym_move_args_t); }

int navgym_policy_features(const float *scan, int n, const float *w1, const float *b1, const float *w2,
                           const float *b2, float *features, void *stream)
{
    if (n <= 0) return 0;
    if (!scan || !w1 || !b1 || !w2 || !b2 || !features) return (int)cudaErrorInvalidValue;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = n < 4 * sms ? n : 4 * sms;  // 4 CTAs of 48 KB shared memory per SM
    policy_features_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(scan, n, w1, b1, w2, b2, features);
    g_launches++;
    return (int)cudaGetLastError();
}

int navgym_peds_move(const navgym_move_args_t *args, void *stream)
{
    const int n = args->num_envs * args->max_ped;
    if (n <= 0) return 0;
    if (!args->mean || !args->v_pref || !args->has_legs || !args->pose || !args->vel || !args->dist_travelled ||
        !args->prev_action || !args->rows)
        return (int)cudaErrorInvalidValue;
    peds_move_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*args);
    g_launches++;
    return (int)cudaGetLastError();
}

int navgym_peds_plan(const navgym_plan_args_t *args, void *stream)
{
    const int n = args->num_envs * args->max_ped;
    if (n <= 0) return 0;
    if (!args->maps || !args->fields || !args->goals || !args->pose || !args->goal_id || !args->waypoint || !args->goal_local)
        return (int)cudaErrorInvalidValue;
    if (args->cand_pose && (!args->free_xy || !args->robot_state || !args->pose_rw || !args->v_pref || !args->has_legs ||
                            !args->dist_travelled || !args->vel || !args->prev_action || !args->cand_v_pref ||
                            !args->cand_legs || !args->cand_goal || !args->cand_rows))
        return (int)cudaErrorInvalidValue;
    peds_plan_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*args);
    g_launches++;
    return (int)cudaGetLastError();
}

int navgym_agent_scan_batch(const navgym_scan_args_t *args, void *stream)
{
    if (args->num_envs <= 0 || args->agents_per_env <= 0) return 0;
    if (args->num_beams <= 0 || !args->pose || !args->lin || !args->ranges) return (int)cudaErrorInvalidValue;
    if (!args->segs && args->robot_state && args->agents_per_env + 1 > 128) return (int)cudaErrorInvalidValue;
    agent_scan_kernel<<<args->num_envs * args->agents_per_env, 128, 0, (cudaStream_t)stream>>>(*args);
    g_launches++;
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

int navgym_raymarching_edt_host(const navgym_raymarching_t *rm, float *out_host)
{
    if (!rm) return (int)cudaErrorInvalidValue;
    return (int)cudaMemcpy(out_host, rm->dist, (size_t)rm->H * rm->W * sizeof(float), cudaMemcpyDeviceToHost);
}

void navgym_raymarching_destroy(navgym_raymarching_t *rm)
{
    if (!rm) return;
    cudaFree(rm->dist);
    delete rm;
}

}  // extern "C"
