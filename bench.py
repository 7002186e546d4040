#!/usr/bin/env python
"""Benchmark of the ProteoScribe sampling hot path (BASELINE.json metric: generated protein
sequences/sec for a full 1024-step decode).

    python bench.py [--gpus N] [--steps K] [--warmup W]          # this repo's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]    # the reference algorithm on host cores

One bench "step" = one full decode (1024 denoising steps, each a full 16-layer forward + draw +
unmask) of one batch of 64 sequences per GPU (BASELINE.json configs[1]; N>1 = configs[2]'s
64-per-GPU sharding with an all-gather of the token ids).  Random-init weights with the reference
key schema and synthetic z_c / paths (no checkpoints or datasets offline): data = "synthetic".

`value`  inputs resident in HBM, C-ABI decode, CUDA events, max over ranks.
`e2e`    the same metric through the drop-in call batch_generate_denoised_sampled() with HOST
         tensors in and the host trajectory out (H2D + D2H inside the timed region).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'generated protein sequences/sec (full 1024-step decode)'
UNIT = 'sequences/s'
BATCH = 64            # sequences per GPU (BASELINE.json configs[1])
L = 1024
C = 29
# algorithmic FLOPs (SURVEY.md section 8d / BASELINE.md section 2)
FLOP_PER_SEQ_STEP = 109.552e9
D = 512


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d['hbm_gbs'], tf_burst=d['bf16_tflops'], tf_sustained=d['bf16_tflops_sustained'], src='measured')
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src='fallback')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index: int):
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ''
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ------------------------------------------------------------------------------------------------
def cpu_oracle_rate(batch: int, max_steps: int, budget_s: float, warm: int = 1):
    """Reference algorithm (oracle port) on the host cores: full forward over all positions, draw at
    every position, cross-sample unmask write, per-step host lists.  Returns (seq/s extrapolated to
    a 1024-step decode, steps timed, seconds, threads)."""
    import torch
    from biom3_b200 import synthetic
    from oracle.model import OracleModel
    from oracle import sampler as osamp
    torch.set_num_threads(os.cpu_count() or 1)
    args = synthetic.stage3_args()
    sd = synthetic.random_state_dict(args, seed=0)
    orc = OracleModel(args, sd)
    z = synthetic.synthetic_z_c(1, 512, seed=1).repeat(batch, 1)
    path = synthetic.synthetic_paths(batch, L, seed=2)
    g = torch.Generator().manual_seed(3)
    n = warm + max_steps
    noise = torch.empty(n, batch * L, C).exponential_(1, generator=g)
    stamps = []

    def hook(i, logits):
        stamps.append(time.perf_counter())

    t_start = time.perf_counter()

    class Budget(Exception):
        pass

    def hook2(i, logits):
        hook(i, logits)
        if i > warm and time.perf_counter() - stamps[warm] > budget_s:
            raise Budget()

    try:
        osamp.decode(orc, torch.zeros(batch, L), torch.zeros(batch).long(), z, path, noise, L, max_iters=n, logits_hook=hook2)
        stamps.append(time.perf_counter())
    except Budget:
        pass
    # stamps[i] = time the forward of iteration i finished; step time = consecutive differences
    timed = stamps[warm:]
    steps = len(timed) - 1
    secs = timed[-1] - timed[0]
    del t_start
    return batch / (secs / steps * L), steps, secs, torch.get_num_threads()


def run_reference_arm(a):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    sample_b, per_step = 8, 2
    rates, secs_all = [], []
    for i in range(a.warmup + a.steps):
        r, steps, secs, threads = cpu_oracle_rate(sample_b, per_step, 1e9, warm=1)
        if i >= a.warmup:
            rates.append(r)
            secs_all.append(secs)
    value = statistics.mean(rates)
    sample = (f'each step = {per_step} consecutive denoising steps of a {sample_b}-sequence batch (full 16-layer fp32 '
              f'forward over all 1024 positions + draw at every position + unmask), extrapolated x1024/{per_step} to a '
              f'full decode; per-step cost is step-independent (same shapes every step)')
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': a.gpus, 'steps': a.steps,
        'warmup': a.warmup, 'ms_per_step': 1e3 * BATCH / value, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'ProteoScribe full-length decode, batch 64 sequences (reference algorithm on host cores, bounded sample)',
                   'global_batch': BATCH, 'seq_len': L, 'denoising_steps': L},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'note': 'reference packages (linear-attention-transformer, axial-positional-embedding) are not installable offline; '
                'this arm times the CPU oracle port of the reference algorithm (oracle/), all host threads',
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def run_cuda_arm(a):
    import torch
    import torch.distributed as dist
    from biom3_b200 import distributed as bdist
    from biom3_b200 import synthetic
    from biom3_b200.Stage3_source import cond_diff_transformer_layer as mod
    from biom3_b200.Stage3_source import sampling_analysis as samp

    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device: biom3_b200 has no CPU path')
    rank, world, local = bdist.init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device(f'cuda:{local}')
    peaks = load_peaks()

    args = synthetic.stage3_args(batch_size_sample=BATCH, num_replicas=BATCH * world)
    args.device = str(dev)
    args.b200_precision = a.precision
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):       # the reference-compatible get_model prints; keep stdout = one JSON line
        model = mod.get_model(args, (32, 32), C)
    model.load_state_dict(synthetic.random_state_dict(args, seed=0))
    model.eval().to(dev)
    eng = model.engine(BATCH)

    # one unit per rank (configs[2]: 512 replicas = 8 units of 64, one per GPU)
    z_host = synthetic.synthetic_z_c(1, 512, seed=1).repeat(BATCH, 1)
    path_host = synthetic.synthetic_paths(BATCH, L, seed=2 + rank)
    z_dev, path_dev = z_host.to(dev), path_host.to(dev)
    stream = torch.cuda.Stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()

    def gather(tokens_u8):
        if world > 1:
            out = torch.empty(world * BATCH, L, dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(out, tokens_u8)
            return out
        return tokens_u8

    def one_decode(seed):
        tokens, _ = eng.decode(z_dev, path_dev, seed=seed)
        return gather(tokens.to(torch.uint8))

    def timed(fn, n):
        """n calls bracketed by barrier + synchronize, CUDA events on the launching stream; max over ranks."""
        barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for i in range(n):
                fn(1000 + i)
            e1.record(stream)
        torch.cuda.synchronize(dev)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    with torch.cuda.stream(stream):
        for i in range(a.warmup):
            one_decode(i)
    torch.cuda.synchronize(dev)

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms_total = timed(one_decode, a.steps)
    clk = clocks.stop() if rank == 0 else None
    value = world * BATCH * a.steps / (ms_total / 1e3)

    # ---- e2e: the drop-in call, host tensors in, host trajectory out
    def one_e2e(seed):
        states, _ = samp.batch_generate_denoised_sampled(
            args=args, model=model, extract_digit_samples=torch.zeros(BATCH, L), extract_time=torch.zeros(BATCH).long(),
            extract_digit_label=z_host, sampling_path=path_host, seed=seed)
        last = states[-1]                                   # what the reference CLI consumes
        if world > 1:
            gather(torch.from_numpy(last[:, 0].astype('uint8')).to(dev))
        return last

    e2e_steps = max(1, min(a.steps, 2))
    with torch.cuda.stream(stream):
        one_e2e(7)
    torch.cuda.synchronize(dev)
    barrier()
    t0 = time.perf_counter()
    with torch.cuda.stream(stream):
        for i in range(e2e_steps):
            one_e2e(2000 + i)
    torch.cuda.synchronize(dev)
    t_e2e = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = world * BATCH * e2e_steps / float(t_e2e.item())
    h2d = BATCH * 512 * 4 + BATCH * L * 8                    # z_c fp32 + sampling_path int64 (state0 is all-mask)
    d2h = L * BATCH * L                                      # uint8 trajectory [T, B, L]

    # ---- per-kernel roofline: CUDA events around every launch of one un-graphed decode step
    profs = [eng.profile_step(BATCH, BATCH) for _ in range(3)]
    prof = {k: statistics.median(p[k] for p in profs) for k in profs[0]}
    prof['launches'], prof['compact_rows'] = int(prof['launches']), int(prof['compact_rows'])
    depth = 16
    M = BATCH * L
    # Last-layer row compaction: the last layer's out-proj / FF1 / FF2 run on `compact_rows` selected token rows instead
    # of M (the rest of that layer's output is never read).  Per-kernel rates count the FLOPs EXECUTED: a category's
    # time covers (depth - 1) full launches + one launch on compact_rows rows.
    mc = int(prof.get('compact_rows', 0))
    full_equiv = depth if mc == 0 else depth - 1 + mc / M           # full-size launches the category's time is worth
    kern = {
        'gemm_ff1 (tcgen05, bias+GELU epilogue)': (prof['gemm_ff1_ms'] / full_equiv, 2.0 * M * 2048 * 512),
        'gemm_ff2 (tcgen05, bias+residual epilogue)': (prof['gemm_ff2_ms'] / full_equiv, 2.0 * M * 512 * 2048),
        'gemm_qkv (tcgen05, head-major store)': (prof['gemm_qkv_ms'] / depth, 2.0 * M * 1536 * 512),
        'gemm_out (tcgen05, bias+residual epilogue)': (prof['gemm_out_ms'] / full_equiv, 2.0 * M * 512 * 512),
        'local_attention': (prof['local_attn_ms'] / depth, 360448.0 * M),
        'linear_attention': (prof['linear_attn_ms'] / depth, 32768.0 * M),
    }
    dom = max(kern, key=lambda k_: kern[k_][0])
    dom_ms, dom_flop = kern[dom]
    achieved = dom_flop / (dom_ms * 1e-3) / 1e12
    executed_flop_per_seq_step = FLOP_PER_SEQ_STEP - (0 if mc == 0 else (1.0 - mc / M) * L * (524288 + 4194304))
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(dom.split(' ')[0])
    roofline = {
        'kernel': dom, 'bound': 'tensor', 'achieved': achieved, 'peak': peaks['tf_sustained'], 'unit': 'TFLOP/s',
        'frac': achieved / peaks['tf_sustained'], 'traffic': traffic,
        'peak_source': f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['src']}; kernel timed inside a long step)",
        'how': 'median of 3 un-graphed decode steps with CUDA events around every launch on the launching stream, right after the timed region',
        'step_ms_ungraphed': prof['total_ms'],
        'share_of_step': dom_ms * (full_equiv if 'attention' not in dom and 'qkv' not in dom else depth) / prof['total_ms'],
        'whole_step_tflops': FLOP_PER_SEQ_STEP * BATCH / (ms_total / a.steps / L * 1e-3) / 1e12,
        'whole_step_frac_of_peak': FLOP_PER_SEQ_STEP * BATCH / (ms_total / a.steps / L * 1e-3) / 1e12 / peaks['tf_sustained'],
        'whole_step_note': 'whole_step_* use the reference algorithm\'s FLOPs (109.552 GFLOP per sequence per step, SURVEY.md 8d); '
                           'whole_step_executed_* count only the FLOPs this build executes (last-layer row compaction skips '
                           'rows whose output the sampler never reads)',
        'last_layer_compact_rows': mc,
        'whole_step_executed_tflops': executed_flop_per_seq_step * BATCH / (ms_total / a.steps / L * 1e-3) / 1e12,
        'whole_step_executed_frac_of_peak': executed_flop_per_seq_step * BATCH / (ms_total / a.steps / L * 1e-3) / 1e12 / peaks['tf_sustained'],
        'per_kernel_ms_per_step': {k_: round(v, 4) for k_, v in prof.items()},
        'per_kernel_tflops': {k_: round(f / (ms * 1e-3) / 1e12, 1) for k_, (ms, f) in kern.items() if ms > 0},
    }

    launches = a.steps * (L * eng.launches_per_step + 9)

    cpu_base = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        v, steps, secs, threads = cpu_oracle_rate(1, 64, 12.0, warm=2)
        cpu_base = {'value': v, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                    'sample': f'{steps} consecutive denoising steps of a 1-sequence batch ({secs:.1f} s of CPU work, fp32 PyTorch oracle '
                              f'of the reference algorithm), extrapolated to 1024 steps'}

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': a.steps, 'warmup': a.warmup,
            'ms_per_step': ms_total / a.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'bf16' if a.precision == 'bf16' else 'f32 (bf16 x 3 split products on the tensor cores, fp32 accumulate / softmax / LayerNorm / GELU)',
            'data': 'synthetic',
            'config': {'workload': f'ProteoScribe full-length decode, batch 64 sequences per GPU, {a.precision} (BASELINE.json configs[1]; '
                                   'N>1: configs[2] sharding, 64 per GPU + NCCL all-gather of token ids)',
                       'global_batch': BATCH * world, 'seq_len': L, 'denoising_steps': L, 'parallelism': f'units{world}',
                       'l2': 'per-step working set (737 MB of activations at B=64) is larger than the 126 MB L2; no flush needed',
                       'noise': 'Exp(1) race noise drawn on the device (Philox) inside the timed region'},
            'ms_per_denoising_step': ms_total / a.steps / L,
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'steps': e2e_steps, 'api': 'biom3_b200.Stage3_source.sampling_analysis.batch_generate_denoised_sampled'},
            'roofline': roofline, 'cpu_baseline': cpu_base, 'clocks': clk, 'gpu_launches': launches,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
def run_c4(a):
    """BASELINE.json configs[3]: Facilitator z_t -> z_c + ProteoScribe sampling for 16 prompts x 32 replicas (units of 32),
    units round-robin over the ranks, two units fused per launch (B = 64, group = 32), token ids all-gathered.  `value`:
    inputs resident, engine calls; `e2e`: the reference-facing path (Facilitator module + batch_stage3_generate_sequences,
    host z_t in, strings out)."""
    import contextlib
    import torch
    import torch.distributed as dist
    from biom3_b200 import distributed as bdist
    from biom3_b200 import run_ProteoScribe_sample as cli
    from biom3_b200 import synthetic
    from biom3_b200.Stage1_source.model import Facilitator
    from biom3_b200.Stage3_source import cond_diff_transformer_layer as mod

    rank, world, local = bdist.init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device(f'cuda:{local}')
    P, R, BS = 16, 32, 32
    args = synthetic.stage3_args(batch_size_sample=BS, num_replicas=R)
    args.device = str(dev)
    args.b200_precision = a.precision
    with contextlib.redirect_stdout(sys.stderr):
        model = mod.get_model(args, (32, 32), C)
    model.load_state_dict(synthetic.random_state_dict(args, seed=0))
    model.eval().to(dev)
    fac = Facilitator(512, 1024, 512, dropout=0.0)
    fac.load_state_dict(synthetic.facilitator_state_dict(512, 1024, seed=31))
    z_t_host = torch.randn(P, 512, generator=torch.Generator().manual_seed(32)) * 1.45
    z_t_dev = z_t_host.to(dev)
    units = bdist.plan_units(P, R, BS)
    mine = bdist.units_for_rank(len(units), rank, world)
    eng = model.engine(2 * BS)
    paths = [synthetic.synthetic_paths(BS, L, seed=100 + u).to(dev) for u in range(len(units))]
    stream = torch.cuda.Stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()

    def one_job(seed0):
        z_c = fac(z_t_dev)                                           # [16, 512] on the device
        local_tok = torch.zeros(len(mine), BS, L, dtype=torch.uint8, device=dev)
        for i in range(0, len(mine), 2):
            batch = mine[i:i + 2]
            z = torch.cat([z_c[units[u][0]].unsqueeze(0).repeat(BS, 1) for u in batch])
            pth = torch.cat([paths[u] for u in batch])
            gs = torch.tensor([seed0 * 1000 + u for u in batch], dtype=torch.int64, device=dev)
            tok, _ = eng.decode(z, pth, group=BS, group_seeds=gs)
            local_tok[i:i + len(batch)] = tok.to(torch.uint8).view(len(batch), BS, L)
        return bdist.gather_unit_tokens(local_tok, mine, len(units), BS)

    with torch.cuda.stream(stream):
        for i in range(a.warmup):
            one_job(i)
    torch.cuda.synchronize(dev)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for i in range(a.steps):
            allt = one_job(10 + i)
        e1.record(stream)
    torch.cuda.synchronize(dev)
    barrier()
    clk = clocks.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = P * R * a.steps / (ms_total / 1e3)
    # e2e: host z_t -> Facilitator -> CLI entry point -> strings (every rank runs the CLI; rank 0 keeps the dictionary)
    barrier()
    t0 = time.perf_counter()
    with torch.cuda.stream(stream):
        torch.manual_seed(77)
        z_c_host = fac(z_t_host)                                     # host in -> host out (H2D + D2H inside)
        seqs = cli.batch_stage3_generate_sequences(args, model, z_c_host)
    torch.cuda.synchronize(dev)
    t_e2e = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    assert len(seqs) == R and all(len(v) == P for v in seqs.values())
    if rank == 0:
        per_launch = L * eng.launches_per_step + 9
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': a.steps, 'warmup': a.warmup,
            'ms_per_step': ms_total / a.steps, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
            'dtype': a.precision, 'data': 'synthetic',
            'config': {'workload': 'Facilitator z_t->z_c + ProteoScribe end-to-end sampling, 16 prompts x 32 replicas '
                                   '(BASELINE.json configs[3]): 16 units of 32, round-robin over the ranks, two units fused per '
                                   'launch (B = 64, group = 32), NCCL all-gather of token ids',
                       'global_batch': P * R, 'seq_len': L, 'denoising_steps': L, 'parallelism': f'units{world}',
                       'l2': 'per-step working set (737 MB of activations at B=64) is larger than the 126 MB L2; no flush needed',
                       'noise': 'Exp(1) race noise drawn on the device (Philox), one seed per unit'},
            'e2e': {'value': P * R / float(t_e2e.item()), 'unit': UNIT, 'h2d_bytes_per_step': P * 512 * 4 + P * 512 * 4 + P * R * L * 8,
                    'd2h_bytes_per_step': P * 512 * 4 + P * R * L, 'steps': 1,
                    'api': 'biom3_b200.Stage1_source.model.Facilitator + biom3_b200.run_ProteoScribe_sample.batch_stage3_generate_sequences'},
            'roofline': None, 'cpu_baseline': None, 'clocks': clk,
            'gpu_launches': a.steps * ((len(mine) + 1) // 2) * per_launch,
            'checksum': int(allt.sum().item()),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_c5(a):
    """BASELINE.json configs[4]: per-step forward microbench, L = 1024, batch sweep 1..1024, fp32-class vs bf16.  One JSON
    line: `value` = sequence-steps per second of the bf16 forward at B = 64; the whole sweep under `sweep`."""
    import torch
    from biom3_b200 import synthetic
    from biom3_b200.engine import Engine
    dev = torch.device('cuda:0')
    torch.cuda.set_device(dev)
    args = synthetic.stage3_args()
    sd = synthetic.random_state_dict(args, seed=0)
    peaks = load_peaks()
    sweep = []
    clocks = ClockSampler(0)
    clocks.start()
    for prec in ('bf16', 'fp32'):
        for B in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024):
            eng = Engine(args, sd, dev, B, precision=prec)
            g = torch.Generator().manual_seed(B)
            x = torch.randint(0, C, (B, L), generator=g).to(dev)
            t = torch.randint(0, L, (B,), generator=g).to(dev)
            zc = synthetic.synthetic_z_c(B, 512, seed=1).to(dev)
            for _ in range(max(3, a.warmup)):
                eng.forward(x, t, zc)
            torch.cuda.synchronize(dev)
            n = max(a.steps, 3) * (8 if B <= 64 else 2)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                eng.forward(x, t, zc)
            e1.record()
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / n
            sweep.append({'precision': prec, 'B': B, 'ms': round(ms, 4), 'seq_steps_per_s': round(B / ms * 1e3, 1),
                          'tflops': round(FLOP_PER_SEQ_STEP * B / (ms * 1e-3) / 1e12, 1)})
            eng.close()
            del eng
            torch.cuda.empty_cache()
    clk = clocks.stop()
    ref = next(r for r in sweep if r['precision'] == 'bf16' and r['B'] == 64)
    line = {
        'metric': 'per-step forward throughput (sequence-steps/s, sequence length 1024)', 'value': ref['seq_steps_per_s'],
        'unit': 'sequence-steps/s', 'n_gpus': 1, 'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': ref['ms'],
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16 (headline) and f32-class (sweep)',
        'data': 'synthetic',
        'config': {'workload': 'ProteoScribe per-step forward microbench: sequence length 1024, batch sweep 1-1024, fp32 vs bf16 '
                               '(BASELINE.json configs[4]); forward API biom3_forward (logits for every position), CUDA-graph replay',
                   'seq_len': L, 'l2': 'B >= 16: activations exceed the 126 MB L2; B < 16 is launch / weight-bandwidth bound'},
        'sweep': sweep, 'peak_tflops_sustained': peaks['tf_sustained'], 'clocks': clk, 'e2e': None, 'roofline': None,
        'cpu_baseline': None, 'gpu_launches': None,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=2)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', type=str, default='cuda', choices=['cuda', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--config', type=str, default='c2', choices=['c2', 'c4', 'c5'],
                    help='c2: BASELINE configs[1] / [2] (default, the driver contract); c4: configs[3] Facilitator + 16 x 32; '
                         'c5: configs[4] forward microbench sweep')
    ap.add_argument('--precision', type=str, default='bf16', choices=['bf16', 'fp32'],
                    help='fp32 = the fp32-class mode (the arithmetic the reference itself runs at inference)')
    a = ap.parse_args()
    # stdout must carry exactly one JSON line: libraries that write to file descriptor 1 behind Python's back (NCCL prints
    # its version banner there) are sent to stderr, and print() is pointed at the saved descriptor.
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)
    sys.stdout = real_stdout
    if a.impl == 'reference':
        run_reference_arm(a)
    else:
        import __graft_entry__
        __graft_entry__.build()
        if a.config == 'c4':
            run_c4(a)
        elif a.config == 'c5':
            run_c5(a)
        else:
            run_cuda_arm(a)
    real_stdout.flush()


if __name__ == '__main__':
    main()
