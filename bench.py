#!/usr/bin/env python
"""bench.py — NavGym-v0 hot-path throughput on B200 (contract: see the task statement).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

A "step" is one lockstep NavGymEnv.step over the whole batch (kinematics -> lidar raycast ->
collision / goal checks -> reward + observation assembly, reference env.py:591-728).  At N=1
the workload is BASELINE.json configs[1]: 4096 batched NavGym-v0 envs on one static indoor
map, no pedestrians, default 512-beam lidar, per-episode scan noise, random actions,
device-side auto-reset from a precomputed spawn pool.  N>1 (torchrun) shards environments:
4096 per GPU, no collective on the step path (weak scaling).

Prints ONE JSON line (rank 0).  `value` = env-steps/s with inputs resident in HBM; `e2e` = the
same through the host-buffer call (pinned actions in, obs/reward/done out, copies timed).
`--impl reference` times the CPU restatement of the reference's path (oracle/, OpenMP over
all host cores) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ENVS_PER_GPU = 4096
NB = 512
METRIC = "env-steps/sec"
BYTES_PER_ENV_STEP = 2197  # obs 2076 + reward 4 + done 1 + info 12 + action 8 + state 96
WORKLOAD = ("NavGym-v0 x%d envs/GPU, static indoor map 1000x1000 @0.05m (corridor_width 3, "
            "iterations 100), no pedestrians, 512-beam 360deg/25m lidar, scan noise "
            "U[0,0.05], random actions, auto-reset from a 65536-tuple spawn pool")


def measured_peak():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md)'


def build_world(seed=0, pool_n=65536):
    from nav_gym_b200 import maps
    rng = np.random.RandomState(seed)
    m = maps.create_indoor_map(3, 100, rng)
    pool = maps.spawn_pool(m, pool_n, rng)
    return m, pool


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                 '-lms', '50'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        rows = [r for (t, r) in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for (_, r) in self.rows]
        for r in rows:
            f = [x.strip() for x in r.split(',')]
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except Exception:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------ CPU arm
def cpu_reference_run(steps, warmup, B_cpu=512, seed=0):
    """The reference's per-step path restated on the CPU (oracle/navgym_oracle.c, OpenMP over
    all host cores; the reference itself is Python + absent native deps and cannot run here).
    Bounded sample: B_cpu envs of the same world and action law."""
    from oracle import oracle as orc
    orc.use_all_cores()
    m, pool = build_world(seed, pool_n=8192)
    rng = np.random.RandomState(seed + 1)
    rows = pool[rng.randint(len(pool), size=B_cpu)]
    o = orc.OracleBatch([m], np.zeros(B_cpu, np.int32), rows[:, 0:2], rows[:, 2:4], rows[:, 4],
                        params=dict(t_stop=502.0))
    sigma = rng.uniform(0, 0.05, B_cpu).astype(np.float32)
    o.reset_obs(want_hits=False)

    def one():
        act = rng.uniform([0, -0.64], [0.5, 0.64], (B_cpu, 2)).astype(np.float32)
        noise = (rng.standard_normal((B_cpu, 2, NB)).astype(np.float32) * sigma[:, None, None])
        t = time.perf_counter()
        o.step(act, noise=noise, want_hits=False)
        dt = time.perf_counter() - t
        d = np.where(o.done)[0]
        if len(d):  # host-side respawn from the pool (the device does this in-kernel)
            r = pool[rng.randint(len(pool), size=len(d))]
            for col, src in ((orc.S_PX, 0), (orc.S_PY, 1), (orc.S_GX, 2), (orc.S_GY, 3), (orc.S_TH, 4),
                             (orc.S_PPX, 0), (orc.S_PPY, 1)):
                o.state[col, d] = r[:, src]
            o.state[orc.S_PV, d] = 0
            o.state[orc.S_PW, d] = 0
            o.steps[d] = 0
        return dt
    for _ in range(warmup):
        one()
    tot = sum(one() for _ in range(steps))
    return dict(value=B_cpu * steps / tot, ms_per_step=1e3 * tot / steps, cores=orc.num_threads(),
                B=B_cpu)


def cpu_baseline_block(sample_steps=20):
    r = cpu_reference_run(sample_steps, 2)
    return {"value": r['value'], "unit": METRIC, "cores": r['cores'], "kind": "port",
            "rays_per_s": r['value'] * NB,
            "sample": "%d envs x %d steps of the same world/action law, C restatement "
                      "(oracle/navgym_oracle.c) with OpenMP over %d threads; noise pre-drawn; the "
                      "reference itself (one env per Python process, pip natives absent here) "
                      "cannot run on this box" % (r['B'], sample_steps, r['cores'])}


# ------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=2000)    # 2000 timed steps
    ap.add_argument('--warmup', type=int, default=200)   # after 200 warm-up steps
    ap.add_argument('--impl', default='native', choices=['native', 'reference'])
    ap.add_argument('--envs', type=int, default=ENVS_PER_GPU, help='environments per GPU')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-flush', action='store_true')
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))

    if a.impl == 'reference' or world == 1:
        # launchers such as torchrun export OMP_NUM_THREADS=1; the CPU legs use every host core
        # (set before libgomp is first loaded: it reads its environment once)
        ncpu = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
        os.environ['OMP_NUM_THREADS'] = str(ncpu)
        os.environ['OMP_PROC_BIND'] = 'false'
        os.environ['OMP_WAIT_POLICY'] = 'active'
    if a.impl == 'reference':
        if rank != 0:
            return
        r = cpu_reference_run(min(a.steps, 400), min(a.warmup, 10))
        line = {"impl": "reference", "metric": METRIC, "value": r['value'], "unit": "env-steps/s",
                "n_gpus": a.gpus, "steps": min(a.steps, 400), "warmup": min(a.warmup, 10),
                "ms_per_step": r['ms_per_step'], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32 scan / f64 pose", "data": "synthetic",
                "rays_per_s": r['value'] * NB,
                "config": {"workload": WORKLOAD % a.envs, "sample_envs": r['B']},
                "cpu_baseline": {"value": r['value'], "unit": "env-steps/s", "cores": r['cores'],
                                 "kind": "port",
                                 "sample": "%d envs per step, oracle C restatement + OpenMP" % r['B']},
                "e2e": {"value": r['value'], "unit": "env-steps/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit('bench.py --impl native needs a CUDA device (no CPU fallback)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    from nav_gym_b200 import _lib
    from nav_gym_b200.batched_env import BatchedNavGym, MapPool, filter_spawn_pool

    B = a.envs
    m, pool = build_world(0)
    pool = filter_spawn_pool(m, pool, dev)
    mp = MapPool([m], dev, spawn_pools=[pool])
    env = BatchedNavGym(B, mp, device=dev, seed=1234, env_offset=rank * B, auto_reset=True)
    env.reset_from_spawn_pool(np.random.RandomState(100 + rank))
    K, W = a.steps, a.warmup
    g = torch.Generator(device=dev)
    g.manual_seed(7 + rank)
    lo = torch.tensor([0.0, -0.64], device=dev)
    hi = torch.tensor([0.5, 0.64], device=dev)
    n_bank = min(K, 64)
    bank = lo + (hi - lo) * torch.rand(n_bank, B, 2, device=dev, generator=g)
    flush = None if a.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    lib = _lib.load()

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # ---- device-resident leg ---------------------------------------------------------
    sampler = ClockSampler(local) if rank == 0 else None
    for i in range(W):
        env.step(bank[i % n_bank])
    if sampler is not None:  # nvidia-smi needs a moment to deliver its first sample
        t_wait = time.time()
        while not sampler.rows and time.time() - t_wait < 3.0:
            time.sleep(0.05)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    launches0 = lib.navgym_launch_count()
    t0 = time.time()
    for i in range(K):
        if flush is not None:
            flush.zero_()
        ev[i][0].record()
        env.step(bank[i % n_bank])
        ev[i][1].record()
    barrier()
    t1 = time.time()
    launches = int(lib.navgym_launch_count() - launches0)
    clocks = sampler.stop(t0, t1) if sampler else None
    step_ms = np.array([s.elapsed_time(e) for s, e in ev])
    total_ms = float(step_ms.sum())
    crash_frac = float(env.is_crash.float().mean().item())
    done_frac = float(env.done.float().mean().item())

    # ---- end-to-end legs: HOST buffers, copies inside the timed region ----------------------
    # (a) synchronous call: navgym_step_batch_host = H2D actions -> step -> D2H obs/reward/done
    # (b) several env groups in flight (submit/wait): each group's next actions are only handed in
    #     after its previous observations have landed on the host; while the host holds group
    #     A's results, group B is stepping, so PCIe hides behind the raycast.
    act_h = torch.empty(n_bank, B, 2, dtype=torch.float32).pin_memory()
    act_h.copy_(bank.cpu())
    obs_h = torch.empty(B, NB + 7, dtype=torch.float32).pin_memory()
    rew_h = torch.empty(B, dtype=torch.float32).pin_memory()
    done_h = torch.empty(B, dtype=torch.uint8).pin_memory()
    Ke = min(K, 1000)
    for i in range(30):
        env.step_host(act_h[i % n_bank], obs_h, rew_h, done_h)
    barrier()
    t_s = time.perf_counter()
    for i in range(Ke):
        env.step_host(act_h[i % n_bank], obs_h, rew_h, done_h)
    barrier()
    e2e_sync_ms = (time.perf_counter() - t_s) * 1e3

    n_groups = int(os.environ.get('NAVGYM_HOST_GROUPS', '4'))
    cur = torch.empty(B, 2, dtype=torch.float32).pin_memory()
    bounds = env.host_groups(n_groups, cur, obs_h, rew_h, done_h)
    cur_np, bank_np = cur.numpy(), act_h.numpy()

    def pipelined(n):
        cur_np[:] = bank_np[0]
        for g in range(n_groups):
            env.submit_host(g)
        for i in range(n):
            nxt = bank_np[(i + 1) % n_bank]
            for g, (b0, b1) in enumerate(bounds):
                env.wait_host(g)                      # group g's obs / reward / done are on the host
                if i + 1 < n:
                    cur_np[b0:b1] = nxt[b0:b1]        # "policy": the group's next actions
                    env.submit_host(g)
    pipelined(60)
    barrier()
    t_s = time.perf_counter()
    pipelined(Ke)
    barrier()
    e2e_ms = (time.perf_counter() - t_s) * 1e3

    # the same bytes with no simulation at all: what this box's PCIe allows for the copy pattern
    streams = [torch.cuda.Stream(device=dev) for _ in bounds]
    torch.cuda.synchronize(dev)
    barrier()  # all ranks copy at the same time, as in the e2e legs
    t_s = time.perf_counter()
    for i in range(Ke):
        for s_, (b0, b1) in zip(streams, bounds):
            with torch.cuda.stream(s_):
                obs_h[b0:b1].copy_(env.obs[b0:b1], non_blocking=True)
    torch.cuda.synchronize(dev)
    barrier()
    copy_only_ms = (time.perf_counter() - t_s) * 1e3

    t = torch.tensor([total_ms, e2e_ms, e2e_sync_ms, copy_only_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, e2e_sync_ms, copy_only_ms = float(t[0]), float(t[1]), float(t[2]), float(t[3])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    kernel_ms = float(step_ms.mean())
    achieved = B * BYTES_PER_ENV_STEP / (kernel_ms * 1e-3) / 1e9
    value = world * B * K / (total_ms * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 scan / f64 pose", "data": "synthetic",
        "rays_per_s": value * NB,
        "config": {"workload": WORKLOAD % B, "envs_per_gpu": B, "global_envs": world * B,
                   "parallelism": "env-shard x%d, no step-path collective" % world,
                   "l2": "flushed between timed steps (256 MiB memset)" if flush is not None else "not flushed",
                   "crash_frac_last_step": crash_frac, "done_frac_last_step": done_frac},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": TRAFFIC_BYTES_PER_LAUNCH,
                     "peak_source": peak_src, "kernel": "step_kernel<false>",
                     "algorithmic_bytes_per_launch": B * BYTES_PER_ENV_STEP,
                     "kernel_ms": kernel_ms,
                     "note": "latency/issue-bound gather kernel: see DESIGN.md roofline section"},
        # second ceiling (SURVEY 8d ii): every march sample is one dependent 4-byte gather from the
        # L2-resident EDT; the peak is the dependent-random-gather rate tools/gather_peak.cu measured
        # on this pool's B200 at the kernel's launch shape (profiles/r1_gather_peak.txt)
        "roofline_gather": {"bound": "l2-gather", "achieved": B * GATHERS_PER_ENV_STEP / (kernel_ms * 1e-3) / 1e9,
                            "peak": GATHER_PEAK_G, "unit": "G gathers/s",
                            "frac": B * GATHERS_PER_ENV_STEP / (kernel_ms * 1e-3) / 1e9 / GATHER_PEAK_G,
                            "gathers_per_env_step": GATHERS_PER_ENV_STEP,
                            "peak_source": "tools/gather_peak.cu on B200, 4 MB table, 64 warps/SM (profiles/r1_gather_peak.txt)"},
        "e2e": {"value": world * B * Ke / (e2e_ms * 1e-3), "unit": "env-steps/s",
                "h2d_bytes_per_step": B * 2 * 4, "d2h_bytes_per_step": B * ((NB + 7) * 4 + 4 + 1),
                "steps": Ke, "timing": "host wall clock, max over ranks",
                "api": ("BatchedNavGym.submit_host/wait_host (C ABI navgym_step_batch_host_submit/"
                        "_wait): pinned host actions in, pinned host obs/reward/done out, %d env "
                        "groups in flight; a group's next actions are submitted only after its "
                        "previous results landed") % n_groups,
                "sync_value": world * B * Ke / (e2e_sync_ms * 1e-3),
                "copy_only_value": world * B * Ke / (copy_only_ms * 1e-3),
                "copy_only_note": "every rank's observation rows copied D2H in the same chunks, all ranks at once, with no stepping: the PCIe / host-memory ceiling of the e2e figure on this box",
                "sync_api": "BatchedNavGym.step_host (navgym_step_batch_host), one blocking call per step"},
        "gpu_launches": launches,
        "clocks": clocks,
    }
    if not a.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline_block()
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# dram__bytes_read.sum + dram__bytes_write.sum of step_kernel<false> per launch, from the
# `ncu --set full` capture summarised in profiles/ (None until measured).
TRAFFIC_BYTES_PER_LAUNCH = 3193600  # profiles/r1_step_kernel_ncu_full_summary.txt
# EDT gathers per env-step on the bench world: 512 beams x 6.94 march samples per ray
# (oracle/analysis/march_stats.py, first sample shared per scan) x 1.01 scans per step.
GATHERS_PER_ENV_STEP = 3590
GATHER_PEAK_G = 700.0  # profiles/r1_gather_peak.txt

if __name__ == '__main__':
    main()
