// her_kernel.cuh -- part of navgym_b200.cu (included there; one translation unit).
// compute_rewards / compute_terminals / compute_info on stored observations (env.py:464-589).
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
