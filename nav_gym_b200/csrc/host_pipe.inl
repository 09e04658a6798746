// host_pipe.inl -- part of navgym_b200.cu (included there; one translation unit).
// Host-buffer entry points (inside extern "C"): navgym_step_batch_host, the asynchronous
// submit / wait form with CUDA-graph replay per env group, and the rollout driver that rotates
// the groups from C with a policy callback (no per-group work in the caller's language).
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
    int chunks, num_envs, device;
    int n_streams;          // streams created so far (create may fail half way)
    cudaStream_t streams[NAVGYM_MAX_CHUNKS];
    cudaEvent_t ready;
    bool have_ready;
    cudaEvent_t landed[NAVGYM_MAX_CHUNKS];  // a group's results are on the host
    int n_landed;
    bool in_flight[NAVGYM_MAX_CHUNKS];
    int32_t *sched[NAVGYM_MAX_CHUNKS];
    int phase[NAVGYM_MAX_CHUNKS];
    int b0[NAVGYM_MAX_CHUNKS + 1];
    navgym_group_graph graphs[NAVGYM_MAX_CHUNKS][3];
    int use_graphs;
};

// The pipe's streams, events and schedule buffers live on the device that was current when it
// was created; every entry point switches to it for the call (and back), so a caller whose
// current device is another one cannot enqueue on foreign-device streams.
struct navgym_device_guard {
    int prev;
    bool switched;
    explicit navgym_device_guard(int dev) : prev(dev), switched(false)
    {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
    }
    ~navgym_device_guard() { if (switched) cudaSetDevice(prev); }
};

void navgym_host_pipe_destroy(navgym_host_pipe_t *p)
{
    if (!p) return;
    navgym_device_guard guard(p->device);
    for (int c = 0; c < p->n_streams; c++) cudaStreamSynchronize(p->streams[c]);
    for (int c = 0; c < p->chunks; c++) {
        for (int i = 0; i < 3; i++)
            if (p->graphs[c][i].exec) cudaGraphExecDestroy(p->graphs[c][i].exec);
        if (p->sched[c]) cudaFree(p->sched[c]);
    }
    for (int c = 0; c < p->n_streams; c++) cudaStreamDestroy(p->streams[c]);
    for (int c = 0; c < p->n_landed; c++) cudaEventDestroy(p->landed[c]);
    if (p->have_ready) cudaEventDestroy(p->ready);
    delete p;
}

navgym_host_pipe_t *navgym_host_pipe_create(int chunks, int num_envs, int longest_first)
{
    if (chunks < 1 || chunks > NAVGYM_MAX_CHUNKS || num_envs < 1) return nullptr;
    navgym_host_pipe_t *p = new navgym_host_pipe_t();  // zero-initialised
    p->chunks = chunks;
    p->num_envs = num_envs;
    if (cudaGetDevice(&p->device) != cudaSuccess) { delete p; return nullptr; }
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);  // hi is the numerically lowest = highest priority
    for (int c = 0; c <= chunks; c++) p->b0[c] = (int)((long long)num_envs * c / chunks);
    // on any failure below: destroy() releases exactly what has been created so far
    for (int c = 0; c < chunks; c++) {
        static const int spread = env_int("NAVGYM_PIPE_PRIO", 1);
        int prio = spread ? hi + c : lo;
        if (prio > lo) prio = lo;
        if (cudaStreamCreateWithPriority(&p->streams[c], cudaStreamNonBlocking, prio) != cudaSuccess) {
            navgym_host_pipe_destroy(p);
            return nullptr;
        }
        p->n_streams = c + 1;
        if (cudaEventCreateWithFlags(&p->landed[c], cudaEventDisableTiming) != cudaSuccess) {
            navgym_host_pipe_destroy(p);
            return nullptr;
        }
        p->n_landed = c + 1;
        if (longest_first) {
            const size_t n = 3 * NAVGYM_SCHED_BUCKETS + (size_t)3 * NAVGYM_SCHED_BUCKETS * num_envs;
            int32_t *h = new int32_t[n]();
            const int cnt = p->b0[c + 1] - p->b0[c];
            h[0] = cnt;
            for (int i = 0; i < cnt; i++) h[3 * NAVGYM_SCHED_BUCKETS + i] = p->b0[c] + i;
            cudaError_t err = cudaMalloc(&p->sched[c], n * sizeof(int32_t));
            if (!err) err = cudaMemcpy(p->sched[c], h, n * sizeof(int32_t), cudaMemcpyHostToDevice);
            delete[] h;
            if (err) { navgym_host_pipe_destroy(p); return nullptr; }
        }
    }
    if (cudaEventCreateWithFlags(&p->ready, cudaEventDisableTiming) != cudaSuccess) {
        navgym_host_pipe_destroy(p);
        return nullptr;
    }
    p->have_ready = true;
    p->use_graphs = env_int("NAVGYM_HOST_GRAPHS", 1);
    return p;
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

// Wait for everything enqueued on the pipe's streams; returns the first error seen.
static int pipe_drain(navgym_host_pipe_t *p)
{
    int first = 0;
    for (int c = 0; c < p->n_streams; c++) {
        const int e = (int)cudaStreamSynchronize(p->streams[c]);
        if (e && !first) first = e;
        p->in_flight[c] = false;
    }
    return first;
}

int navgym_step_batch_host(navgym_host_pipe_t *p, const navgym_step_args_t *args, void *stream,
                           const float *actions_host, float *obs_host, float *reward_host,
                           uint8_t *done_host)
{
    if (!p || args->num_envs != p->num_envs || !args->actions) return (int)cudaErrorInvalidValue;
    navgym_device_guard guard(p->device);
    cudaStream_t in = (cudaStream_t)stream;
    const size_t B = (size_t)args->num_envs;
    CK(cudaMemcpyAsync((void *)args->actions, actions_host, B * 2 * sizeof(float), cudaMemcpyHostToDevice, in));
    CK(cudaEventRecord(p->ready, in));
    int err = 0;
    for (int c = 0; c < p->chunks && !err; c++) {
        cudaStream_t st = p->streams[c];
        navgym_step_args_t a = *args;
        a.env_begin = p->b0[c];
        a.env_count = p->b0[c + 1] - p->b0[c];
        if (a.env_count <= 0) continue;
        a.sched = p->sched[c];
        a.sched_phase = p->phase[c];
        err = (int)cudaStreamWaitEvent(st, p->ready, 0);
        if (!err) err = enqueue_group(a, st, actions_host, obs_host, reward_host, done_host, false);
        if (!err && p->sched[c]) p->phase[c] = (p->phase[c] + 1) % 3;
    }
    // also on error: never return with chunks still in flight on the caller's buffers
    const int drained = pipe_drain(p);
    return err ? err : drained;
}

// Asynchronous variant for callers that keep several groups of environments in flight
// (group g = the pipe's g-th env range): submit enqueues H2D(actions) -> step -> D2H(results)
// for one group on that group's stream and returns at once; wait blocks until that group's
// results have landed.  While the host consumes group A's observations, group B is stepping.
static int submit_group(navgym_host_pipe_t *p, const navgym_step_args_t *args, int group,
                        const float *actions_host, float *obs_host, float *reward_host, uint8_t *done_host)
{
    cudaStream_t st = p->streams[group];
    navgym_step_args_t a = *args;
    a.env_begin = p->b0[group];
    a.env_count = p->b0[group + 1] - p->b0[group];
    if (a.env_count <= 0) return 0;
    a.sched = p->sched[group];
    a.sched_phase = p->phase[group];
    if (p->sched[group]) p->phase[group] = (p->phase[group] + 1) % 3;
    p->in_flight[group] = true;
    if (!p->use_graphs) {
        int err = enqueue_group(a, st, actions_host, obs_host, reward_host, done_host, true);
        if (!err) err = (int)cudaEventRecord(p->landed[group], st);
        return err;
    }
    navgym_group_graph &g = p->graphs[group][a.sched_phase];
    const void *host[4] = {actions_host, obs_host, reward_host, done_host};
    if (!g.exec || memcmp(&g.key, &a, sizeof(a)) != 0 || memcmp(g.host, host, sizeof(host)) != 0) {
        if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
        // pageable host memory cannot be captured: such callers keep the call-by-call path
        for (int i = 0; i < 4; i++)
            if (!is_pinned_host(host[i])) {
                int err = enqueue_group(a, st, actions_host, obs_host, reward_host, done_host, true);
                if (!err) err = (int)cudaEventRecord(p->landed[group], st);
                return err;
            }
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
    return (int)cudaEventRecord(p->landed[group], st);
}

int navgym_step_batch_host_submit(navgym_host_pipe_t *p, const navgym_step_args_t *args, int group,
                                  const float *actions_host, float *obs_host, float *reward_host,
                                  uint8_t *done_host)
{
    if (!p || group < 0 || group >= p->chunks || args->num_envs != p->num_envs || !args->actions)
        return (int)cudaErrorInvalidValue;
    navgym_device_guard guard(p->device);
    return submit_group(p, args, group, actions_host, obs_host, reward_host, done_host);
}

int navgym_step_batch_host_wait(navgym_host_pipe_t *p, int group)
{
    if (!p || group < 0 || group >= p->chunks) return (int)cudaErrorInvalidValue;
    navgym_device_guard guard(p->device);
    p->in_flight[group] = false;
    return (int)cudaStreamSynchronize(p->streams[group]);
}

int navgym_host_pipe_groups(const navgym_host_pipe_t *p) { return p ? p->chunks : 0; }

int navgym_host_pipe_group_bounds(const navgym_host_pipe_t *p, int group, int *begin, int *end)
{
    if (!p || group < 0 || group >= p->chunks) return (int)cudaErrorInvalidValue;
    *begin = p->b0[group];
    *end = p->b0[group + 1];
    return 0;
}

// Built-in policy for benchmarks and tests: step s takes its actions from row (s mod rows) of a
// host action bank f32 [rows][num_envs][2].
void navgym_policy_action_bank(void *user, int group, int env_begin, int env_end, int64_t step,
                               const float *obs_host, const float *reward_host, const uint8_t *done_host,
                               float *actions_host)
{
    (void)group; (void)obs_host; (void)reward_host; (void)done_host;
    const navgym_action_bank_t *b = (const navgym_action_bank_t *)user;
    const float *src = b->actions + ((size_t)(step % b->rows) * b->num_envs + env_begin) * 2;
    memcpy(actions_host + 2 * (size_t)env_begin, src, (size_t)(env_end - env_begin) * 2 * sizeof(float));
}

// The rollout loop in C.  All groups are primed with the policy's step-0 actions; then, `steps`
// times round, group g is waited for (its observations / rewards / dones of step s are on the
// host), the policy writes the group's next actions from them, and the group is submitted again
// -- while the host handles group g, the other groups are stepping or copying.  Groups are
// served round-robin, which keeps every group at the same step count.
int navgym_host_rollout(navgym_host_pipe_t *p, const navgym_step_args_t *args, int64_t steps,
                        navgym_policy_fn policy, void *user, float *actions_host, float *obs_host,
                        float *reward_host, uint8_t *done_host)
{
    if (!p || !policy || steps < 0 || args->num_envs != p->num_envs || !args->actions)
        return (int)cudaErrorInvalidValue;
    if (steps == 0) return 0;
    navgym_device_guard guard(p->device);
    int err = 0;
    for (int g = 0; g < p->chunks && !err; g++) {
        policy(user, g, p->b0[g], p->b0[g + 1], 0, obs_host, reward_host, done_host, actions_host);
        err = submit_group(p, args, g, actions_host, obs_host, reward_host, done_host);
    }
    for (int64_t s = 0; s < steps && !err; s++) {
        for (int g = 0; g < p->chunks && !err; g++) {
            err = (int)cudaEventSynchronize(p->landed[g]);
            p->in_flight[g] = false;
            if (err || s + 1 >= steps) continue;
            policy(user, g, p->b0[g], p->b0[g + 1], s + 1, obs_host, reward_host, done_host, actions_host);
            err = submit_group(p, args, g, actions_host, obs_host, reward_host, done_host);
        }
    }
    const int drained = pipe_drain(p);
    return err ? err : drained;
}
