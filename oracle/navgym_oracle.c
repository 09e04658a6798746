/*
 * navgym_oracle.c — CPU ORACLE for the NavGym-v0 per-step hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is product code: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library, and only as the checker / the timed CPU baseline.  The product
 * (nav_gym_b200/) never links, imports or executes it.
 *
 * What it restates (citations relative to /root/reference/nav_gym/src/nav_gym_env/):
 *   - KetiRobot.set_vel                      keti_robot.py:64-93
 *   - NavGymEnv._compute_scan                env.py:385-441
 *   - xy_to_ij / batch_xy_to_ij              env.py:1228-1258
 *   - _convert_obs / angle_correction        env.py:443-462, utils.py:5-9
 *   - compute_rewards/terminals/info         env.py:464-589
 *   - step() ordering and crash rollback     env.py:591-728
 * and the THIRD-PARTY natives the reference calls but does not vendor
 * (nav_gym/setup.py:23-26, un-pinned, sources absent from /root/reference):
 *   - range_libc  PyOMap / PyRayMarching.calc_range_many   (call site env.py:337-340,425)
 *     published algorithm: exact Euclidean distance transform (Felzenszwalb &
 *     Huttenlocher lower-envelope passes) + sphere tracing with step max(0.999 d, 1).
 *   - pymap2d     flatten_contours / render_contours_in_lidar / render_agents_in_lidar /
 *     CSimAgent   (call site env.py:14,398-402,428-432): analytic ray-segment and
 *     ray-disc nearest hit, min-merged into the scan.
 *
 * PARITY STATUS.  First-party arithmetic (kinematics, cell mapping, obs, reward,
 * terminals, rollback) is PINNED: tests/golden/ holds traces produced by executing the
 * reference's own env.py unmodified in the build container (oracle/make_golden.py) and
 * this file reproduces them.  The third-party natives are "PARITY UNPINNED": their
 * sources and any golden vectors are absent, so the semantics below are the canonical
 * ones this project defines (DESIGN.md "Canonical native semantics"); the CUDA path is
 * held bit-exact against THEM.
 *
 * Floating-point contract: build with -O2 -ffp-contract=off -fno-fast-math so that
 * every + - * / sqrt below is one IEEE-754 rounding, and fmaf()/fma() are the only
 * fused operations.  The CUDA kernels spell the same sequence with __f*_rn intrinsics.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define NVO_INF_G 32768 /* column distance when a column holds no occupied cell */
#define NVO_NB 512      /* beams of the robot lidar, keti_robot.py:48 */

/* ---------------------------------------------------------------- EDT ------------- */
/* Exact squared Euclidean distance transform, 0 at occupied cells, cell units.
 * Pass 1: per column, integer distance to the nearest occupied cell in that column.
 * Pass 2: per row, lower envelope of parabolas (Felzenszwalb-Huttenlocher 1-D dt).
 * occ/dist are [H][W] row-major, row = y (map_info['data'][y][x]). */
static void dt1d(const double *f, int n, double *d, int *v, double *z)
{
    int k = 0;
    v[0] = 0;
    z[0] = -1e300;
    z[1] = 1e300;
    for (int q = 1; q < n; q++) {
        double s;
        for (;;) {
            int p = v[k];
            s = ((f[q] + (double)q * q) - (f[p] + (double)p * p)) / (2.0 * q - 2.0 * p);
            if (s <= z[k] && k > 0)
                k--;
            else
                break;
        }
        if (s <= z[k]) { /* k == 0 and q dominates everywhere left of it */
            v[0] = q;
            z[0] = -1e300;
            z[1] = 1e300;
        } else {
            k++;
            v[k] = q;
            z[k] = s;
            z[k + 1] = 1e300;
        }
    }
    k = 0;
    for (int q = 0; q < n; q++) {
        while (z[k + 1] < (double)q)
            k++;
        double dq = (double)(q - v[k]);
        d[q] = dq * dq + f[v[k]];
    }
}

void nvo_edt_sq(const uint8_t *occ, int H, int W, int32_t *d2)
{
    int32_t *g = (int32_t *)malloc(sizeof(int32_t) * (size_t)H * W);
    for (int x = 0; x < W; x++) {
        int last = -1;
        for (int y = 0; y < H; y++) {
            if (occ[(size_t)y * W + x])
                last = y;
            g[(size_t)y * W + x] = last < 0 ? NVO_INF_G : y - last;
        }
        last = -1;
        for (int y = H - 1; y >= 0; y--) {
            if (occ[(size_t)y * W + x])
                last = y;
            int32_t dn = last < 0 ? NVO_INF_G : last - y;
            if (dn < g[(size_t)y * W + x])
                g[(size_t)y * W + x] = dn;
        }
    }
    double *f = (double *)malloc(sizeof(double) * W);
    double *d = (double *)malloc(sizeof(double) * W);
    double *z = (double *)malloc(sizeof(double) * (W + 1));
    int *v = (int *)malloc(sizeof(int) * W);
    for (int y = 0; y < H; y++) {
        for (int x = 0; x < W; x++) {
            double gv = (double)g[(size_t)y * W + x];
            f[x] = gv * gv;
        }
        dt1d(f, W, d, v, z);
        for (int x = 0; x < W; x++)
            d2[(size_t)y * W + x] = (int32_t)d[x];
    }
    free(f); free(d); free(z); free(v); free(g);
}

/* float EDT as range_libc's DistanceTransform holds it: sqrtf of the squared map. */
void nvo_edt(const uint8_t *occ, int H, int W, float *dist)
{
    int32_t *d2 = (int32_t *)malloc(sizeof(int32_t) * (size_t)H * W);
    nvo_edt_sq(occ, H, W, d2);
    for (size_t i = 0; i < (size_t)H * W; i++)
        dist[i] = sqrtf((float)d2[i]);
    free(d2);
}

/* ------------------------------------------------------- beam direction ----------- */
/* cos / sin of a beam heading.  range_libc calls cosf/sinf; to make the
 * direction a function of the heading bits alone — independent of which libm or GPU math
 * library evaluates it — the canonical contract spells it out: Cody-Waite
 * reduction by pi/2 in two fma steps, the fdlibm kernel polynomials in Horner/fma form,
 * quadrant fix-up, result rounded to float.  Max error 1.8e-16 before that rounding; on 2e7
 * random float headings in [-12, 18] the float result equalled correctly rounded cosf/sinf
 * every time (oracle/README: sincos check). */
static const double NV_TWO_OVER_PI = 6.36619772367581382433e-01;
static const double NV_PIO2_HI = 1.57079632679489655800e+00, NV_PIO2_LO = 6.12323399573676603587e-17;
static const double NV_S1 = -1.66666666666666324348e-01, NV_S2 = 8.33333333332248946124e-03,
                    NV_S3 = -1.98412698298579493134e-04, NV_S4 = 2.75573137070700676789e-06,
                    NV_S5 = -2.50507602534068634195e-08, NV_S6 = 1.58969099521155010221e-10;
static const double NV_C1 = 4.16666666666666019037e-02, NV_C2 = -1.38888888888741095749e-03,
                    NV_C3 = 2.48015872894767294178e-05, NV_C4 = -2.75573143513906633035e-07,
                    NV_C5 = 2.08757232129817482790e-09, NV_C6 = -1.13596475577881948265e-11;

void nvo_sincos(double x, double *sn, double *cs)
{
    double k = rint(x * NV_TWO_OVER_PI);
    double r = fma(-k, NV_PIO2_HI, x);
    r = fma(-k, NV_PIO2_LO, r);
    double z = r * r;
    double ps = fma(z, NV_S6, NV_S5);
    ps = fma(z, ps, NV_S4); ps = fma(z, ps, NV_S3); ps = fma(z, ps, NV_S2); ps = fma(z, ps, NV_S1);
    double s = fma(r * z, ps, r);
    double pc = fma(z, NV_C6, NV_C5);
    pc = fma(z, pc, NV_C4); pc = fma(z, pc, NV_C3); pc = fma(z, pc, NV_C2); pc = fma(z, pc, NV_C1);
    double c = fma(z * z, pc, fma(z, -0.5, 1.0));
    int n = (int)k & 3;
    double a = (n & 1) ? c : s, b = (n & 1) ? s : c;
    *sn = (n & 2) ? -a : a;
    *cs = ((n + 1) & 2) ? -b : b;
}

/* ------------------------------------------------------------ ray marching -------- */
/* range_libc RayMarching::calc_range restated, canonical form:
 * (dx, dy) = float(nvo_sincos(h)); sample cell = trunc(fmaf(dx, t, x0)); occupied <=> d <= 0;
 * t += max(d * 0.999f, 1.0f); out of map or t >= t_stop -> max_range.
 * hit[0..1] receives (px - x0, py - y0) as integers, or (INT16_MIN, INT16_MIN). */
/* x0 + dx * t: fused by default; -DNVO_MARCH_NO_FMA = separately rounded multiply and add (the
 * real range_libc is built with -ffast-math, its source is absent: either may be what its
 * binary does -- the CUDA library has the matching -DNAVGYM_MARCH_NO_FMA build). */
#ifdef NVO_MARCH_NO_FMA
#define NVO_MARCH_POS(d, t, o) ((d) * (t) + (o))
#else
#define NVO_MARCH_POS(d, t, o) fmaf((d), (t), (o))
#endif
int nvo_march_is_fused(void)
{
#ifdef NVO_MARCH_NO_FMA
    return 0;
#else
    return 1;
#endif
}
float nvo_calc_range(const float *dist, int W, int H, float x0, float y0, float heading,
                     float max_range, float t_stop, int32_t *hit, int32_t *nsteps)
{
    double sn, cs;
    nvo_sincos((double)heading, &sn, &cs);
    float dx = (float)cs, dy = (float)sn;
    float t = 0.0f;
    int32_t n = 0;
    if (hit) { hit[0] = INT16_MIN; hit[1] = INT16_MIN; }
    while (t < t_stop) {
        int px = (int)NVO_MARCH_POS(dx, t, x0);
        int py = (int)NVO_MARCH_POS(dy, t, y0);
        if (px < 0 || px >= W || py < 0 || py >= H)
            break;
        float d = dist[(size_t)py * W + px];
        n++;
        if (d <= 0.0f) {
            float xd = (float)px - x0;
            float yd = (float)py - y0;
            if (hit) { hit[0] = (int32_t)xd; hit[1] = (int32_t)yd; }
            if (nsteps) *nsteps = n;
            return sqrtf(xd * xd + yd * yd);
        }
        float st = d * 0.999f;
        t = t + (st > 1.0f ? st : 1.0f);
    }
    if (nsteps) *nsteps = n;
    return max_range;
}

/* PyRayMarching.calc_range_many(ins f32[N,3], outs f32[N]) — env.py:425 */
void nvo_calc_range_many(const float *dist, int W, int H, const float *ins, float *outs,
                         int N, float max_range, float t_stop, int32_t *hits,
                         int32_t *nsteps)
{
    for (int i = 0; i < N; i++)
        outs[i] = nvo_calc_range(dist, W, H, ins[3 * i], ins[3 * i + 1], ins[3 * i + 2],
                                 max_range, t_stop, hits ? hits + 2 * i : 0,
                                 nsteps ? nsteps + i : 0);
}

/* ------------------------------------------------ contour / disc rendering -------- */
/* Canonical float32 ray-segment hit: ray o + t d, segment a + u (b - a).
 * Returns t >= 0 or +inf. */
static inline float seg_hit(float ox, float oy, float dx, float dy, float ax, float ay,
                            float bx, float by)
{
    float ex = bx - ax, ey = by - ay;
    float wx = ax - ox, wy = ay - oy;
    float den = dx * ey - dy * ex;
    if (den == 0.0f)
        return INFINITY;
    float tn = wx * ey - wy * ex;
    float un = wx * dy - wy * dx;
    float t = tn / den;
    float u = un / den;
    if (t >= 0.0f && u >= 0.0f && u <= 1.0f)
        return t;
    return INFINITY;
}

/* Canonical float32 ray-disc nearest non-negative root. */
static inline float disc_hit(float ox, float oy, float dx, float dy, float X, float Y,
                             float r)
{
    float cx = X - ox, cy = Y - oy;
    float b = dx * cx + dy * cy;
    float c = (cx * cx + cy * cy) - r * r;
    float q = b * b - c;
    if (q < 0.0f)
        return INFINITY;
    float s = sqrtf(q);
    float t = b - s;
    if (t < 0.0f)
        t = b + s;
    if (t < 0.0f)
        return INFINITY;
    return t;
}

/* render_contours_in_lidar(ranges inout, angles, flat[V,3]=(contour id,x,y), lidar_xy)
 * env.py:430-431.  Every contour is closed (last vertex -> first).  dirs holds the beam
 * directions (dx,dy) as float32 pairs. */
void nvo_render_contours(float *ranges, const float *dirs, int K, const float *flat, int V,
                         const float *lidar_xy)
{
    float ox = lidar_xy[0], oy = lidar_xy[1];
    int s = 0;
    while (s < V) {
        int e = s;
        while (e + 1 < V && flat[3 * (e + 1)] == flat[3 * s])
            e++;
        for (int v = s; v <= e; v++) {
            int w = (v == e) ? s : v + 1;
            float ax = flat[3 * v + 1], ay = flat[3 * v + 2];
            float bx = flat[3 * w + 1], by = flat[3 * w + 2];
            for (int k = 0; k < K; k++) {
                float t = seg_hit(ox, oy, dirs[2 * k], dirs[2 * k + 1], ax, ay, bx, by);
                if (t < ranges[k])
                    ranges[k] = t;
            }
        }
        s = e + 1;
    }
}

/* segments given directly as [S,4] = (ax,ay,bx,by) */
void nvo_render_segments(float *ranges, const float *dirs, int K, const float *segs, int S,
                         const float *lidar_xy)
{
    float ox = lidar_xy[0], oy = lidar_xy[1];
    for (int s = 0; s < S; s++)
        for (int k = 0; k < K; k++) {
            float t = seg_hit(ox, oy, dirs[2 * k], dirs[2 * k + 1], segs[4 * s],
                              segs[4 * s + 1], segs[4 * s + 2], segs[4 * s + 3]);
            if (t < ranges[k])
                ranges[k] = t;
        }
}

/* CMap2D.render_agents_in_lidar with the agents already reduced to discs [D,3]=(x,y,r)
 * env.py:432 */
void nvo_render_discs(float *ranges, const float *dirs, int K, const float *discs, int D,
                      const float *lidar_xy)
{
    float ox = lidar_xy[0], oy = lidar_xy[1];
    for (int d = 0; d < D; d++)
        for (int k = 0; k < K; k++) {
            float t = disc_hit(ox, oy, dirs[2 * k], dirs[2 * k + 1], discs[3 * d],
                               discs[3 * d + 1], discs[3 * d + 2]);
            if (t < ranges[k])
                ranges[k] = t;
        }
}

/* beam directions: heading_k = (float)(lin[k] + (double)theta32) (env.py:388-390,
 * 420-424); dirs = nvo_sincos of that float heading, rounded to float. */
void nvo_beam_dirs(const double *lin, int K, float theta32, float *headings, float *dirs)
{
    for (int k = 0; k < K; k++) {
        float h = (float)(lin[k] + (double)theta32);
        if (headings) headings[k] = h;
        double sn, cs;
        nvo_sincos((double)h, &sn, &cs);
        dirs[2 * k] = (float)cs;
        dirs[2 * k + 1] = (float)sn;
    }
}

/* -------------------------------------------------------------- cell mapping ------ */
/* batch_xy_to_ij for one coordinate (env.py:1235-1253).  rule 0 ("numpy1", the
 * reference's pinned NumPy 1.x era: np.float32 scalar / python float evaluates in
 * float64, is stored to a float32 array, then truncated); rule 1 ("numpy2": NEP-50,
 * the division itself is float32 — what the reference does when executed under the
 * build container's NumPy 2.3). */
int32_t nvo_xy_to_cell(float x32, double origin, double res, int dim, int rule)
{
    float c;
    if (rule == 0)
        c = (float)(((double)x32 - origin) / res);
    else
        c = (x32 - (float)origin) / (float)res;
    if (c >= (float)dim) c = (float)(dim - 1);
    if (c < 0.0f) c = 0.0f;
    return (int32_t)c;
}

/* ------------------------------------------------------------- batched step ------- */
typedef struct {
    int32_t W, H;
    int64_t offset; /* into the float EDT pool */
    double ox, oy, res;
} nvo_map_t;

/* state64 is [NS][B] (structure of arrays), rows: */
enum { S_PX, S_PY, S_TH, S_GX, S_GY, S_PPX, S_PPY, S_PYAW, S_PV, S_PW, S_NS };

typedef struct {
    /* constants */
    double dt;                  /* time_step                      __init__.py:8  */
    double dist_thresh;         /* distance_threshold             __init__.py:10 */
    double min_turn_radius;     /* min_turning_radius             __init__.py:9  */
    double r_scale, r_success, r_crash, r_progress, r_forward, r_rotation, r_discomfort;
    float range_max;            /* keti_robot.py:47 */
    float t_stop;               /* cells; reference marches to W*H (env.py:337) */
    int32_t cell_rule;
    int32_t max_disc, max_seg;
    int32_t num_scan_stack;     /* S of env.py:257-279; obs row = S*512 + 7 */
} nvo_params_t;

static void scan_once(const nvo_params_t *P, const nvo_map_t *m, const float *edt,
                      const double *lin, double px, double py, double th,
                      const float *discs, int nd, const float *segs, int ns,
                      const float *noise, float *scan, int16_t *hits)
{
    float lx = (float)px, ly = (float)py, lt = (float)th; /* env.py:386 */
    float dirs[2 * NVO_NB], head[NVO_NB];
    nvo_beam_dirs(lin, NVO_NB, lt, head, dirs);
    int32_t ci = nvo_xy_to_cell(lx, m->ox, m->res, m->H, P->cell_rule);
    int32_t cj = nvo_xy_to_cell(ly, m->oy, m->res, m->W, P->cell_rule);
    float max_range = (float)((double)m->W * (double)m->H);
    float res32 = (float)m->res;
    for (int k = 0; k < NVO_NB; k++) {
        int32_t hit[2];
        float r = nvo_calc_range(edt + m->offset, m->W, m->H, (float)ci, (float)cj, head[k],
                                 max_range, P->t_stop, hit, 0);
        scan[k] = r * res32; /* env.py:426 */
        if (hits) { hits[2 * k] = (int16_t)hit[0]; hits[2 * k + 1] = (int16_t)hit[1]; }
    }
    float lxy[2] = {lx, ly};
    if (ns) nvo_render_segments(scan, dirs, NVO_NB, segs, ns, lxy);
    if (nd) nvo_render_discs(scan, dirs, NVO_NB, discs, nd, lxy);
    for (int k = 0; k < NVO_NB; k++) { /* env.py:435-440 */
        float r = scan[k];
        if (r < 0.0f) r = 0.0f;
        if (r > P->range_max) r = P->range_max;
        if (noise && r != P->range_max) r = r + noise[k];
        scan[k] = r;
    }
}

/* _stack_scan (env.py:257-279): [current scan x (S-1-n) | the n previous scans, oldest first |
 * current scan]; hist holds the scans of the last n <= S-1 returned observations. */
static void stack_row(float *o, const float *scan, int S, const float *hist, int n)
{
    int idx = 0;
    for (int p = 0; p < S - 1 - n; p++) memcpy(o + (size_t)NVO_NB * idx++, scan, sizeof(float) * NVO_NB);
    for (int h = 0; h < n; h++) memcpy(o + (size_t)NVO_NB * idx++, hist + (size_t)h * NVO_NB, sizeof(float) * NVO_NB);
    memcpy(o + (size_t)NVO_NB * (S - 1), scan, sizeof(float) * NVO_NB);
}

/* prev_obs_queue.append(obs), a deque(maxlen = S-1) (env.py:727, 738) */
static void push_hist(float *hist, int32_t *n, int S, const float *scan)
{
    if (S <= 1) return;
    if (*n == S - 1) {
        memmove(hist, hist + NVO_NB, sizeof(float) * NVO_NB * (size_t)(S - 2));
        *n -= 1;
    }
    memcpy(hist + (size_t)NVO_NB * *n, scan, sizeof(float) * NVO_NB);
    *n += 1;
}

/* One lockstep NavGymEnv.step over B environments.
 *  actions f32[B,2]; discs f32[B,max_disc,3], ndisc i32[B]; segs f32[B,max_seg,4], nseg;
 *  noise f32[B,2,512] or NULL (slot 0: the step's scan, slot 1: the crash re-scan);
 *  obs f32[B,519]; tail64 f64[B,7]; reward f64[B]; done u8[B]; is_success u8[B];
 *  is_crash u8[B]; distance f64[B]; hits i16[B,512,2] or NULL (first scan). */
void nvo_step_batch(const nvo_params_t *P, int B, const nvo_map_t *maps, const float *edt,
                    const int32_t *map_id, const double *lin, const float *thr,
                    const float *dthr, double *state, int32_t *steps, const float *actions,
                    const float *discs, const int32_t *ndisc, const float *segs,
                    const int32_t *nseg, const float *noise, float *obs, double *tail64,
                    double *reward, uint8_t *done, uint8_t *is_success, uint8_t *is_crash,
                    double *distance, int16_t *hits, float *hist, int32_t *nhist)
{
    const int SS = P->num_scan_stack > 1 ? P->num_scan_stack : 1;
#pragma omp parallel for schedule(dynamic, 16)
    for (int e = 0; e < B; e++) {
        double *S = state;
#define ST(f) S[(size_t)(f) * B + e]
        const nvo_map_t *m = &maps[map_id[e]];
        double v = (double)actions[2 * e], w = (double)actions[2 * e + 1];
        steps[e] += 1; /* env.py:592 */
        if (P->min_turn_radius > 0) { /* env.py:595-600 */
            double lim = fabs(w) * P->min_turn_radius;
            if (v >= 0) v = v > lim ? v : lim;
            else v = v < -lim ? v : -lim;
        }
        /* keti_robot.py:64-93 */
        double th0 = ST(S_TH);
        double rx = 0.14474 * cos(th0) + ST(S_PX);
        double ry = 0.14474 * sin(th0) + ST(S_PY);
        double th1 = th0 + w * P->dt;
        rx = rx + cos(th1) * v * P->dt;
        ry = ry + sin(th1) * v * P->dt;
        double px = -0.14474 * cos(th1) + rx;
        double py = -0.14474 * sin(th1) + ry;
        double thn = fmod(th1, 2 * M_PI);
        if (thn != 0 && thn < 0) thn += 2 * M_PI;
        ST(S_PX) = px; ST(S_PY) = py; ST(S_TH) = thn;

        float scan[NVO_NB];
        const float *dd = discs ? discs + (size_t)e * P->max_disc * 3 : 0;
        const float *ss = segs ? segs + (size_t)e * P->max_seg * 4 : 0;
        int nd = ndisc ? ndisc[e] : 0, ns = nseg ? nseg[e] : 0;
        scan_once(P, m, edt, lin, px, py, thn, dd, nd, ss, ns,
                  noise ? noise + (size_t)e * 2 * NVO_NB : 0, scan,
                  hits ? hits + (size_t)e * 2 * NVO_NB : 0);
        double yaw = atan2(sin(thn), cos(thn)); /* utils.py:5-9 */
        double ppx = ST(S_PPX), ppy = ST(S_PPY), gx = ST(S_GX), gy = ST(S_GY);
        double pv = ST(S_PV), pw = ST(S_PW);

        /* compute_rewards / compute_terminals / compute_info, env.py:464-589 */
        double dxg = gx - px, dyg = gy - py;
        double dist = sqrt(dxg * dxg + dyg * dyg);
        double dxp = gx - ppx, dyp = gy - ppy;
        double pdist = sqrt(dxp * dxp + dyp * dyp);
        int success = dist < P->dist_thresh;
        int crash = 0, discomf = 0;
        for (int k = 0; k < NVO_NB; k++) {
            if ((double)scan[k] - (double)thr[k] < 0) crash = 1;
            if ((double)scan[k] - (double)dthr[k] < 0) discomf = 1;
        }
        discomf = discomf && !crash;
        double r_s = success ? 1.0 * P->r_success * P->r_scale : 0.0;
        double r_c = crash ? -1.0 * P->r_crash * P->r_scale : 0.0;
        double r_p = (pdist - dist) * P->r_progress * P->r_scale;
        double r_f = pv * P->r_forward * P->r_scale;
        double r_r = -1.0 * (pw * pw) * P->r_rotation * P->r_scale;
        double r_d = 0.0;
        if (discomf) {
            double mn = INFINITY;
            for (int k = 0; k < NVO_NB; k++) {
                float den = (dthr[k] - thr[k]) + 1e-6f; /* float32 array arithmetic */
                double q = ((double)scan[k] - (double)thr[k]) / (double)den;
                if (q < mn) mn = q;
            }
            r_d = -(1.0 - mn) * P->r_discomfort * P->r_scale;
        }
        reward[e] = r_s + r_c + r_p + r_f + r_r + r_d;
        done[e] = (uint8_t)(success || crash);
        is_success[e] = (uint8_t)success;
        is_crash[e] = (uint8_t)crash;
        distance[e] = dist;

        double opx = px, opy = py; /* pose fields of the returned observation */
        if (crash) { /* env.py:707-723 */
            px = ppx; py = ppy; thn = ST(S_PYAW);
            ST(S_PX) = px; ST(S_PY) = py; ST(S_TH) = thn;
            scan_once(P, m, edt, lin, px, py, thn, dd, nd, ss, ns,
                      noise ? noise + (size_t)e * 2 * NVO_NB + NVO_NB : 0, scan, 0);
            yaw = atan2(sin(thn), cos(thn));
            opx = px; opy = py;
        }
        float *o = obs + (size_t)e * (SS * NVO_NB + 7);
        float *he = hist ? hist + (size_t)e * (SS - 1) * NVO_NB : 0;
        stack_row(o, scan, SS, he, nhist ? nhist[e] : 0);
        if (nhist) push_hist(he, &nhist[e], SS, scan);
        double *t7 = tail64 + (size_t)e * 7;
        t7[0] = ppx; t7[1] = ppy; t7[2] = opx; t7[3] = opy; t7[4] = pv; t7[5] = pw; t7[6] = yaw;
        for (int i = 0; i < 7; i++) o[SS * NVO_NB + i] = (float)t7[i];
        /* env.py:725-727 */
        ST(S_PV) = (double)actions[2 * e];
        ST(S_PW) = (double)actions[2 * e + 1];
        if (P->min_turn_radius > 0) ST(S_PV) = v; /* the reference clamps `action` in place */
        ST(S_PPX) = opx; ST(S_PPY) = opy; ST(S_PYAW) = yaw;
#undef ST
    }
}

/* first observation of an episode (reset contract, env.py:822-831): prev_pose = pose,
 * vel = prev_action = 0; noise slot 0 */
void nvo_reset_obs_batch(const nvo_params_t *P, int B, const nvo_map_t *maps, const float *edt,
                         const int32_t *map_id, const double *lin, double *state,
                         int32_t *steps, const float *discs, const int32_t *ndisc,
                         const float *segs, const int32_t *nseg, const float *noise,
                         float *obs, double *tail64, int16_t *hits, float *hist, int32_t *nhist)
{
    const int SS = P->num_scan_stack > 1 ? P->num_scan_stack : 1;
#pragma omp parallel for schedule(dynamic, 16)
    for (int e = 0; e < B; e++) {
        double *S = state;
#define ST(f) S[(size_t)(f) * B + e]
        const nvo_map_t *m = &maps[map_id[e]];
        float scan[NVO_NB];
        const float *dd = discs ? discs + (size_t)e * P->max_disc * 3 : 0;
        const float *ss = segs ? segs + (size_t)e * P->max_seg * 4 : 0;
        int nd = ndisc ? ndisc[e] : 0, ns = nseg ? nseg[e] : 0;
        double px = ST(S_PX), py = ST(S_PY), th = ST(S_TH);
        steps[e] = 0;
        scan_once(P, m, edt, lin, px, py, th, dd, nd, ss, ns,
                  noise ? noise + (size_t)e * 2 * NVO_NB : 0, scan,
                  hits ? hits + (size_t)e * 2 * NVO_NB : 0);
        double yaw = atan2(sin(th), cos(th));
        float *o = obs + (size_t)e * (SS * NVO_NB + 7);
        float *he = hist ? hist + (size_t)e * (SS - 1) * NVO_NB : 0;
        if (nhist) nhist[e] = 0;
        stack_row(o, scan, SS, he, 0);
        if (nhist) push_hist(he, &nhist[e], SS, scan);
        double *t7 = tail64 + (size_t)e * 7;
        t7[0] = px; t7[1] = py; t7[2] = px; t7[3] = py; t7[4] = 0; t7[5] = 0; t7[6] = yaw;
        for (int i = 0; i < 7; i++) o[SS * NVO_NB + i] = (float)t7[i];
        ST(S_PV) = 0; ST(S_PW) = 0; ST(S_PPX) = px; ST(S_PPY) = py; ST(S_PYAW) = yaw;
#undef ST
    }
}

/* ------------------------------------------------ reset path: grid BFS ----------------
 * 4-connected geodesic distance in cells over blocked[H][W] (non-zero = blocked), -1 where
 * unreachable: the uniform-cost form of pyastar2d.astar_path on the 0.25 m cost map
 * (env.py:343-354), used when minting spawn pools (oracle/make_bench_world.py). */
void nvo_grid_bfs(const uint8_t *blocked, int H, int W, int sr, int sc, int32_t *dist)
{
    size_t n = (size_t)H * W;
    for (size_t i = 0; i < n; i++) dist[i] = -1;
    if (sr < 0 || sc < 0 || sr >= H || sc >= W || blocked[(size_t)sr * W + sc]) return;
    int32_t *queue = (int32_t *)malloc(n * sizeof(int32_t));
    size_t head = 0, tail = 0;
    queue[tail++] = sr * W + sc;
    dist[(size_t)sr * W + sc] = 0;
    while (head < tail) {
        int32_t cur = queue[head++];
        int r = cur / W, c = cur % W;
        const int nr[4] = {r + 1, r - 1, r, r}, nc[4] = {c, c, c + 1, c - 1};
        for (int k = 0; k < 4; k++) {
            if (nr[k] < 0 || nc[k] < 0 || nr[k] >= H || nc[k] >= W) continue;
            size_t ni = (size_t)nr[k] * W + nc[k];
            if (blocked[ni] || dist[ni] >= 0) continue;
            dist[ni] = dist[cur] + 1;
            queue[tail++] = (int32_t)ni;
        }
    }
    free(queue);
}

void nvo_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int nvo_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
