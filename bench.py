#!/usr/bin/env python
"""Benchmark of the SCD naming round (BASELINE.json metric: ms per naming round = one k-means iteration
+ full-vocabulary scoring with per-image top-k + per-cluster vote).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config C2]

* default arm: the B200 path.  `value` = ms per round with all inputs resident in HBM (CUDA events, max
  over ranks); `e2e` = the same round through the public Python API with HOST (pinned) buffers, H2D/D2H
  inside the timed region; `roofline` = the dominant kernel (fused scoring/top-k, tensor-core bound)
  against MEASURED_PEAKS.json; `cpu_baseline` = the oracle (a port of the reference's PyTorch path) timed on
  this box's host cores on a bounded row sample (N=1 / rank 0 only).
* N > 1 (launched by torch.distributed.run, one rank per GPU): STRONG scaling of the same workload - rows
  block-sharded for k-means (one packed NCCL all-reduce per iteration), vocabulary column-sharded for naming
  (local fused top-k, all-gather, k-way merge), vote replicated.
* --impl reference: the reference's own CPU implementation of the path (the oracle port - the reference is
  Python/PyTorch, there is nothing to compile into oracle/_ref) with all host threads, on a bounded sample of
  the same workload, extrapolated linearly in rows.  Rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

METRIC = 'ms per naming round (k-means iter + vocab top-k + vote)'
CPU_SAMPLE_ROWS = int(os.environ.get('SCD_BENCH_CPU_ROWS', '8192'))      # rows of the bounded CPU sample (tests shrink it)


# ------------------------------------------------------------------------------------------ helpers
def load_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p.get('hbm_gbs', 6650.0), tflops=p.get('bf16_tflops', 1590.0),
                    tflops_sustained=p.get('bf16_tflops_sustained', 1400.0), source='measured')
    return dict(hbm_gbs=6650.0, tflops=1590.0, tflops_sustained=1400.0, source='fallback')


class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU while the timed region runs: a background
    `nvidia-smi -lms 100` process (the profiling recipe's clocks line) - a separate process, so the
    queries never contend with this process's launch path (in-process NVML polling did: it doubled the
    measured step time)."""

    FIELDS = ['clocks.sm', 'clocks.max.sm', 'power.draw', 'clocks_event_reasons.hw_slowdown',
              'clocks_event_reasons.hw_thermal_slowdown', 'clocks_event_reasons.sw_thermal_slowdown',
              'clocks_event_reasons.sw_power_cap']

    def __init__(self, index, enabled=True):
        self.proc, self.t_start, self.t_stop = None, None, None
        if not enabled:
            return
        try:
            import subprocess
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=timestamp,' + ','.join(self.FIELDS),
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.35)           # let it print its first line before the region starts
        except Exception:
            self.proc = None

    def start(self):
        self.t_start = time.time()

    def result(self):
        self.t_stop = time.time()
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            text, _ = self.proc.communicate(timeout=2)
        except Exception:
            self.proc.kill()
            return out
        import datetime
        sm, reasons, mx, power = [], set(), None, []
        for line in text.strip().splitlines():
            parts = [x.strip() for x in line.split(',')]
            if len(parts) != 1 + len(self.FIELDS):
                continue
            try:
                ts = datetime.datetime.strptime(parts[0], '%Y/%m/%d %H:%M:%S.%f').timestamp()
                clk, mxc = float(parts[1]), float(parts[2])
            except Exception:
                continue
            mx = mxc
            if self.t_start - 0.05 <= ts <= self.t_stop + 0.05:
                sm.append(clk)
                try:
                    power.append(float(parts[3]))
                except Exception:
                    pass
                for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), parts[4:]):
                    if val.lower().startswith('active'):
                        reasons.add(name)
        out.update(sm_mhz=(float(np.median(sm)) if sm else None), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm),
                   power_w_max=(max(power) if power else None))
        return out


def physical_gpu_index(local_rank):
    vis = os.environ.get('CUDA_VISIBLE_DEVICES')
    if vis:
        try:
            return int(vis.split(',')[local_rank])
        except Exception:
            return local_rank
    return local_rank


# ------------------------------------------------------------------------------------------ CPU legs
_CPU_DATA = {}


def cpu_round_ms(cfg, sample_rows, threads, repeats=2):
    """One naming round of the reference's CPU path (oracle port) on `sample_rows` rows of the workload,
    extrapolated linearly to cfg.n rows.  Returns (ms_per_round_full, detail dict)."""
    from oracle import kmeans_oracle, naming_oracle
    from scd_b200 import synth
    torch.set_num_threads(threads)
    n = min(sample_rows, cfg.n)
    if (cfg.name, n) not in _CPU_DATA:
        _CPU_DATA[(cfg.name, n)] = synth.make(cfg, n_rows=n)
    data = _CPU_DATA[(cfg.name, n)]
    X, Xc, W, C0 = data['X'], data['Xc'], data['W'], data['C0']
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        labels, _, _ = kmeans_oracle.estep(X, C0, 1024)                     # pairwise_distance(X, C, 1024) + torch.min
        kmeans_oracle.mstep(X, labels, C0.clone())                          # the :61-64 loop
        t1 = time.perf_counter()
        idx, _ = naming_oracle.score_topk(Xc, W, 5, variant='ptsup')        # main_ptsup.py:526-545 restated
        t2 = time.perf_counter()
        c2c = naming_oracle.vote(idx, labels.numpy(), list(range(cfg.k)), 5)
        naming_oracle.voted_candidates(c2c, list(range(cfg.k)), 20)
        t3 = time.perf_counter()
        cur = dict(kmeans=t1 - t0, naming=t2 - t1, vote=t3 - t2, total=t3 - t0)
        if best is None or cur['total'] < best['total']:
            best = cur
    scale = cfg.n / n
    detail = {k: round(v * 1e3 * scale, 1) for k, v in best.items()}
    return best['total'] * 1e3 * scale, detail, n


# ------------------------------------------------------------------------------------------ GPU arm
class Round:
    """Device-resident state of one rank and the launches of one naming round.

    N = 1: everything local.  N > 1: image rows block-sharded over the ranks for BOTH contractions - k-means
    (one packed NCCL all-reduce of [K*D sums | K counts | inertia] per iteration) and naming (every rank
    holds the whole vocabulary, 32 MB, so its rows' top-k needs no exchange) - then the per-row results
    (labels, top-k name indices; N*48 bytes in total) are all-gathered and the vote runs replicated.
    `naming_shard='vocab'` switches naming to the vocabulary-column-sharded scheme (all rows x V/N columns per
    rank, all-gather of [N, k] lists, k-way merge kernel)."""

    def __init__(self, cfg, rank, world, group, naming_shard='rows'):
        from scd_b200 import dist as sdist, kmeans, naming, synth
        self.cfg, self.rank, self.world, self.group = cfg, rank, world, group
        self.kmeans, self.naming, self.sdist = kmeans, naming, sdist
        self.naming_shard = naming_shard if world > 1 else 'rows'
        data = synth.make(cfg)
        self.host = data
        dev = torch.device('cuda')
        self.row_lo, self.row_hi = sdist.shard_bounds(cfg.n, world, rank)
        n_local = self.row_hi - self.row_lo
        self.X = data['X'][self.row_lo:self.row_hi].to(dev).contiguous()
        self.C = data['C0'].to(dev).contiguous()
        self.C_new = torch.empty_like(self.C)
        self.inertia = torch.zeros(1, dtype=torch.float64, device=dev)
        self.mstep = kmeans._MStep(n_local, synth.D, cfg.k, dev)
        self.km = kmeans.K_Means(k=cfg.k, process_group=group if world > 1 else None)
        if world > 1:
            self.g_labels = sdist.RowGather(cfg.n, (), torch.int64, dev, group)
            self.labels = self.g_labels.local
        else:
            self.g_labels = None
            self.labels = torch.empty(n_local, dtype=torch.int64, device=dev)
        if self.naming_shard == 'rows':
            self.col_lo, self.col_hi = 0, cfg.v
            self.Xc = naming._feats_bf16(data['Xc'][self.row_lo:self.row_hi])
            self.vocab = naming.Vocabulary(data['W'].to(dev))
            if world > 1:
                self.g_idx = sdist.RowGather(cfg.n, (5,), torch.int64, dev, group)
                self.topk = naming.TopKPlan(n_local, cfg.v, 5, dev, idx_out=self.g_idx.local)
            else:
                self.g_idx = None
                self.topk = naming.TopKPlan(n_local, cfg.v, 5, dev)
            self.name_flops = 2.0 * n_local * cfg.v * synth.D
        else:
            self.col_lo, self.col_hi = sdist.shard_bounds(cfg.v, world, rank)
            self.Xc = naming._feats_bf16(data['Xc'])
            self.vocab = naming.Vocabulary(data['W'][:, self.col_lo:self.col_hi].to(dev), col_offset=self.col_lo)
            self.g_idx = None
            self.topk = naming.TopKPlan(cfg.n, self.col_hi - self.col_lo, 5, dev, want_stats=True)
            self.name_flops = 2.0 * cfg.n * (self.col_hi - self.col_lo) * synth.D
        self.vote_plan = naming.VotePlan(cfg.n, cfg.k, 20, dev)
        self.launches_per_round = 0
        self.ev_name = None
        self.graph = None

    def run(self, ev_name=None):
        cfg, km, nm = self.cfg, self.kmeans, self.naming
        launches = 0
        # ---- k-means iteration: E-step, M-step sums, (all-reduce), divide + centre shift
        self.inertia.zero_()
        km._estep(self.X, self.C, self.labels, self.inertia); launches += 2           # centroid split + E-step
        self.mstep.sums_counts(self.X, self.labels); launches += 4                    # hist, scan, scatter, segment sum
        counts_f = self.km._allreduce(self.mstep, self.inertia)
        launches += 1 if counts_f is not None else 0                                  # pack (the all-reduce is NCCL's)
        self.mstep.finalize(self.C, self.C_new, counts_f); launches += 2
        # ---- full-vocabulary scoring + per-image top-5
        if ev_name is not None:
            ev_name[0].record()
        if self.naming_shard == 'rows':
            vals, idx, _, _ = self.topk.run(self.Xc, self.vocab, False); launches += 2
        else:
            vals, idx = self.sdist.sharded_score_topk(self.Xc, self.vocab, 5, False, self.group, plan=self.topk); launches += 3
        if ev_name is not None:
            ev_name[1].record()
        # ---- per-cluster vote over all rows
        if self.world == 1:
            out = nm.vote_device(idx, None, cfg.k, 5, 20, plan=self.vote_plan, presorted=self.mstep); launches += 1
        else:
            labels_all = self.g_labels.gather()
            idx_all = self.g_idx.gather() if self.g_idx is not None else idx
            out = nm.vote_device(idx_all, labels_all, cfg.k, 5, 20, plan=self.vote_plan); launches += 4
        self.launches_per_round = launches
        self.last = (vals, idx, out)
        return out

    def capture(self):
        """Record one round into a CUDA graph (launch-bound at N > 1: ~15 launches of 5-400 us each)."""
        try:
            g = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self.run()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=side):
                self.run()
            self.graph = g
        except Exception as e:                                  # eager launches are always available
            sys.stderr.write(f'[bench] CUDA graph capture failed, timing eager launches: {e}\n')
            self.graph = None
            torch.cuda.synchronize()
        return self.graph is not None

    def step(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            self.run()


def e2e_round(cfg, host, vocab, pinned):
    """One round through the public API from HOST buffers: H2D of the step's inputs, the round, D2H of the results."""
    from scd_b200 import kmeans, naming
    X = pinned['X'].to('cuda', non_blocking=True)
    C = pinned['C0'].to('cuda', non_blocking=True)
    Xc = pinned['Xc']                       # stays on the host: score_topk uploads it in chunks under the kernel
    km = kmeans.K_Means(k=cfg.k, max_iterations=1, n_init=1)
    labels = torch.empty(cfg.n, dtype=torch.int64, device='cuda')
    best_labels, inertia, centers, _ = km._lloyd(X, X, labels, 0, C)
    vals, idx = naming.score_topk(Xc, vocab, k=5, softmax=False)
    names, counts, distinct, rows, ovf = naming.vote_device(idx, best_labels, cfg.k, 5, 20)
    res = [t.to('cpu', non_blocking=True) for t in (best_labels, centers, vals, idx, names, counts)]
    torch.cuda.synchronize()
    h2d = X.numel() * 4 + C.numel() * 4 + Xc.numel() * 4
    d2h = sum(t.numel() * t.element_size() for t in res) + 16
    return h2d, d2h


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', default='C2')
    ap.add_argument('--naming-shard', default='rows', choices=['rows', 'vocab'],
                    help='N > 1: how the scoring/top-k is partitioned (rows: no exchange; vocab: all-gather + k-way merge)')
    ap.add_argument('--no-graph', action='store_true', help='time eager launches instead of CUDA-graph replays')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-clocks', action='store_true')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup

    from scd_b200 import synth
    cfg = synth.CONFIGS[args.config]
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    feat_mb = cfg.n * synth.D * 6 / 1e6
    voc_mb = cfg.v * synth.D * 2 / 1e6
    if world == 1:
        par = 'single GPU'
    elif args.naming_shard == 'rows':
        par = (f'rows/{world} for k-means (packed NCCL all-reduce of sums+counts+inertia) and for naming (vocabulary replicated, '
               f'no exchange); labels + top-k indices all-gathered, vote replicated')
    else:
        par = f'rows/{world} (k-means, all-reduce) x vocab/{world} (naming, all-gather + k-way merge); vote replicated'
    config = dict(workload=f'{cfg.name}: {cfg.n}x{synth.D} image features, K={cfg.k}, V={cfg.v} names, top-5, vote top-20',
                  n=cfg.n, d=synth.D, k=cfg.k, v=cfg.v, topk=5,
                  l2=f'inputs ({feat_mb:.0f} MB features + {voc_mb:.0f} MB vocabulary) exceed the 126 MB L2', parallelism=par)

    # -------------------------------------------------------------------------------- reference arm
    if args.impl == 'reference':
        if rank != 0:
            return 0
        threads = os.cpu_count() or 1
        vals = []
        for s in range(max(args.steps, 1) + max(args.warmup, 0)):
            ms, detail, n_s = cpu_round_ms(cfg, CPU_SAMPLE_ROWS, threads, repeats=1)
            if s >= args.warmup:
                vals.append(ms)
        v = float(np.median(vals))
        sample = f'{n_s} of {cfg.n} rows (all {cfg.k} centroids, all {cfg.v} names), time scaled linearly in rows'
        print(json.dumps(dict(impl='reference', metric=METRIC, value=round(v, 1), unit='ms', n_gpus=args.gpus, steps=args.steps,
                              warmup=args.warmup, ms_per_step=round(v, 1), higher_is_better=False, scaling='strong',
                              vs_baseline=None, dtype='f32', data='synthetic', config=config,
                              cpu_baseline=dict(value=round(v, 1), unit='ms', cores=threads, kind='port', sample=sample, split_ms=detail),
                              e2e=dict(value=round(v, 1), unit='ms', h2d_bytes_per_step=0, d2h_bytes_per_step=0))))
        return 0

    # -------------------------------------------------------------------------------- B200 arm
    assert torch.cuda.is_available(), 'bench.py needs a GPU (the B200 path has no CPU fallback)'
    # stdout carries exactly ONE JSON line: NCCL prints its version banner on fd 1 when NCCL_DEBUG is set (seen on the
    # 2-GPU box), so everything libraries write to fd 1 is sent to stderr and the line goes to the saved descriptor
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)
    torch.cuda.set_device(local_rank)
    group = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
        group = dist.group.WORLD
    peaks = load_peaks()
    rnd = Round(cfg, rank, world, group, args.naming_shard)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier(group=group)
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        rnd.run()
    barrier()
    graphed = False if args.no_graph else rnd.capture()
    if world > 1:                                   # every rank must take the same path (collectives inside)
        import torch.distributed as dist
        flag = torch.tensor([1 if graphed else 0], device='cuda')
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if not bool(flag.item()):
            rnd.graph, graphed = None, False
    if graphed:
        for _ in range(2):
            rnd.step()
    barrier()
    sampler = ClockSampler(physical_gpu_index(local_rank), enabled=not args.no_clocks)
    barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        rnd.step()
    ev1.record()
    barrier()
    clocks = sampler.result()
    ms_total = ev0.elapsed_time(ev1)

    # the dominant kernel, timed live with events around its launch (eager launches, same stream, same inputs)
    n_k = max(3, min(args.steps, 20))
    name_evs = []
    for _ in range(n_k):
        evs = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        rnd.run(ev_name=evs)
        name_evs.append(evs)
    barrier()
    name_ms = float(np.mean([a.elapsed_time(b) for a, b in name_evs]))
    t = torch.tensor([ms_total, name_ms], dtype=torch.float64, device='cuda')
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    ms_per_step = float(t[0]) / args.steps
    name_ms = float(t[1])

    # roofline of the dominant kernel: fused scoring/top-k (tensor-core bound); algorithmic flops = 2 * rows * cols * D of this rank
    flops = rnd.name_flops
    achieved = flops / (name_ms * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'roofline_traffic.json')
    if os.path.exists(tpath) and world == 1 and cfg.name == 'C2':      # the ncu capture is of the single-GPU C2 launch
        with open(tpath) as f:
            traffic = json.load(f).get('name_topk_kernel_dram_bytes_per_launch')
    roofline = dict(bound='tensor', kernel='name_topk_kernel<5> (+ topk_merge)', achieved=round(achieved, 1), peak=peaks['tflops'],
                    unit='TFLOP/s', frac=round(achieved / peaks['tflops'], 4), traffic=traffic,
                    peak_source=f"{peaks['source']} bf16 burst (MEASURED_PEAKS.json); sustained figure: {peaks['tflops_sustained']}",
                    kernel_ms=round(name_ms, 4), flops_per_launch=flops)

    e2e = None
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_e2e:
        pinned = {k: rnd.host[k].pin_memory() for k in ('X', 'Xc', 'C0')}
        for _ in range(2):
            h2d, d2h = e2e_round(cfg, rnd.host, rnd.vocab, pinned)
        torch.cuda.synchronize()
        n_e2e = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            h2d, d2h = e2e_round(cfg, rnd.host, rnd.vocab, pinned)
        torch.cuda.synchronize()
        e2e = dict(value=round((time.perf_counter() - t0) * 1e3 / n_e2e, 3), unit='ms', h2d_bytes_per_step=int(h2d),
                   d2h_bytes_per_step=int(d2h), api='K_Means._lloyd(1 iter) + naming.score_topk (chunked upload under the kernel) + naming.vote_device from pinned host tensors')
    elif world > 1:
        e2e = dict(value=None, unit='ms', h2d_bytes_per_step=0, d2h_bytes_per_step=0, note='measured at N=1 only')
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, detail, n_s = cpu_round_ms(cfg, CPU_SAMPLE_ROWS, threads, repeats=2)
        cpu_baseline = dict(value=round(v, 1), unit='ms', cores=threads, kind='port', split_ms=detail,
                            sample=f'{n_s} of {cfg.n} rows (all {cfg.k} centroids, all {cfg.v} names), time scaled linearly in rows')

    if rank == 0:
        line = dict(metric=METRIC, value=round(ms_per_step, 4), unit='ms', n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=round(ms_per_step, 4), higher_is_better=False, scaling='strong', vs_baseline=None, dtype='bf16',
                    data='synthetic', config=config, clocks=clocks, e2e=e2e, gpu_launches=rnd.launches_per_round * args.steps,
                    launch_mode='cuda-graph replay' if graphed else 'eager',
                    roofline=roofline, cpu_baseline=cpu_baseline,
                    split=dict(naming_ms=round(name_ms, 4), rest_ms=round(ms_per_step - name_ms, 4)))
        real_stdout.write(json.dumps(line) + '\n')
        real_stdout.flush()
    if world > 1:
        # Leave without tearing NCCL down: destroy_process_group() / interpreter exit can block for minutes while
        # a CUDA graph that captured NCCL kernels is still alive (seen on the 2-GPU box).  Every rank has passed
        # the final barrier and rank 0 has flushed its line, so a hard exit loses nothing.
        rnd.graph = None
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    return 0


if __name__ == '__main__':
    sys.exit(main())
