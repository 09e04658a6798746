// native_kernels.cuh -- part of navgym_b200.cu (included there; one translation unit).
// range_libc / pymap2d stand-ins: exact EDT build, calc_range_many, render_*_in_lidar.
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
