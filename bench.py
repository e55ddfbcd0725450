#!/usr/bin/env python
"""Benchmark of the batched decode path (lattice build + feature scoring + beam search).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c2] [--impl reference]

One step = one pass of the hot path over one batch of synthetic sentences (SURVEY.md §8d
generators).  At N=1 the workload is BASELINE.json configs[1]: 10k synthetic Hangul sentences
(avg 20 syllables), 100k-morpheme synthetic dictionary, beam=5.  Under torchrun every rank tags
its own batch (weak scaling, no collective on the data path); the step time is the max over ranks.

Printed JSON line (rank 0): `value` = sentences/s with inputs resident in HBM (CUDA events),
`e2e` = the same through `lt_tag_batch_host` with pinned host buffers (copies inside the timed
region), `roofline` = algorithmic bytes of the dominant kernel / its device time / measured HBM
peak, `cpu_baseline` = the oracle's pure-Python port of the reference timed on this box's cores.
`--impl reference` times that CPU path alone, with all host cores.
"""

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'sentences_per_sec'
UNIT = 'sentences/s'


# ------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------
def measured_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured'
    except Exception:
        return 6650.0, 'fallback'


class ClockSampler:
    """Samples SM clocks and throttle reasons with nvidia-smi while the timed region runs."""

    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
             'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.QUERY, '--format=csv,noheader,nounits',
                 '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.lines:
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(smax) if smax else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU reference arm (oracle port) — the only place bench.py touches oracle/
# ------------------------------------------------------------------------------------------------
_W = {}


def _oracle_objects(config, feature_items, coef_bytes):
    import lattice_based_tagger_b200 as pkg
    from lattice_based_tagger_b200 import synth
    from oracle import lattice_oracle as lo
    cfg = dict(synth.CONFIGS[config])
    alphabet = synth.make_alphabet(cfg['alphabet'], seed=0)
    tag_to_morphs = synth.make_dictionary(cfg['n_dict'], alphabet, seed=0)
    rules = synth.make_rules(tag_to_morphs, alphabet, n_keys=min(5000, max(50, cfg['n_dict'] // 20)), seed=1)
    dictionary = pkg.dictionary.MorphemeDictionary(tag_to_morphs, rules)
    funcs = [pkg.beam.RegularizationScore()]
    if feature_items is not None:
        feature_dic = {tuple(k): i for i, k in enumerate(feature_items)}
        coef = np.frombuffer(coef_bytes, dtype=np.float64)
        funcs.append(pkg.beam.SimpleTrigramFeatureScore(pkg.features.SimpleTrigramEncoder(feature_dic), coef))
    return lo, lo.OracleTagger(dictionary, pkg.beam.BeamScoreFunctions(*funcs))


def _worker_init(config, feature_items, coef_bytes, beam):
    lo, tagger = _oracle_objects(config, feature_items, coef_bytes)
    _W['lo'], _W['tagger'], _W['beam'] = lo, tagger, beam


def _worker_tag(sents):
    lo, tagger, beam = _W['lo'], _W['tagger'], _W['beam']
    counters = lo.Counters()
    out = []
    for sent in sents:
        try:
            best = tagger.tag(sent, beam, counters)
            out.append(([(w[6], w[7], w[3], w[1]) for w in best.words[1:-1]], best.score))
        except IndexError:
            out.append(None)
    return out, counters.as_dict()


class CpuArm:
    """multiprocessing.Pool over all host cores running the oracle's Python port."""

    def __init__(self, config, feature_items, coef, beam, cores=None):
        import multiprocessing as mp
        self.cores = cores or os.cpu_count() or 1
        ctx = mp.get_context('spawn')
        coef_bytes = None if coef is None else np.asarray(coef, dtype=np.float64).tobytes()
        self.pool = ctx.Pool(self.cores, initializer=_worker_init,
                             initargs=(config, feature_items, coef_bytes, beam))
        # make sure every worker finished building its tables before anything is timed
        self.pool.map(_worker_tag, [[] for _ in range(self.cores * 2)])

    def run(self, sents):
        """-> (seconds, results, counters) for tagging `sents` across the pool."""
        chunk = max(1, min(64, len(sents) // (self.cores * 4) or 1))
        chunks = [sents[i:i + chunk] for i in range(0, len(sents), chunk)]
        t0 = time.perf_counter()
        parts = self.pool.map(_worker_tag, chunks)
        dt = time.perf_counter() - t0
        results = [r for part, _ in parts for r in part]
        totals = {}
        for _, c in parts:
            for k, v in c.items():
                totals[k] = totals.get(k, 0) + v
        return dt, results, totals

    def close(self):
        self.pool.close()
        self.pool.join()


def oracle_feature_inputs(config, sample):
    """Feature dictionary for the reference arm, drawn exactly as the GPU arm draws it but with the
    oracle as lattice / best-path provider (the two providers return identical results)."""
    import lattice_based_tagger_b200 as pkg
    from lattice_based_tagger_b200 import synth
    lo, tagger = _oracle_objects(config, None, None)

    class _Seq:
        def __init__(self, hyp):
            self.sequences = [pkg.dictionary.Word(*w) for w in hyp.words]

    def tag_fn(sents):
        out = []
        for s in sents:
            try:
                out.append(_Seq(tagger.tag(s, synth.CONFIGS[config]['beam'])))
            except IndexError:
                out.append(None)
        return out

    def lattice_fn(sents):
        out = []
        for s in sents:
            edges = sorted(tagger.lattice(s), key=lambda w: (w[7], w[6]))     # device order
            words = [pkg.dictionary.Word(*w) for w in edges]
            bindex = []
            if words:
                bindex = [[] for _ in range(len(s.replace(' ', '')))]
                for w in words:
                    bindex[w.b].append(w)
            out.append((words, bindex))
        return out

    cfg = synth.CONFIGS[config]
    tags = list(tagger.view.tag_to_morphs)
    vocab = None
    if cfg['n_feat'] > 1_000_000:
        vocab = [(m, t) for t, ms in tagger.view.tag_to_morphs.items() for m in sorted(ms)]
    return synth.make_features(sample, tag_fn, lattice_fn, cfg['n_feat'], tags, seed=3, vocab=vocab)


def feature_sample(sents):
    return sents[:max(64, min(2000, len(sents) // 10))]


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port), all cores."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from lattice_based_tagger_b200 import synth
    cfg, dictionary, sents = synth.build_workload(args.config, rank=0, n_sent=args.sentences)
    # sentences without any dictionary edge raise in the reference: replace them, as the GPU arm does
    lo, probe = _oracle_objects(args.config, None, None)
    good = [i for i, s in enumerate(sents) if probe.lattice(s)]
    for i in range(len(sents)):
        if not probe.lattice(sents[i]):
            sents[i] = sents[good[i % len(good)]]
    feature_dic, coef = oracle_feature_inputs(args.config, feature_sample(sents))
    items = list(feature_dic.keys())
    arm = CpuArm(args.config, items, coef, cfg['beam'])
    # bounded sample per step so that the whole run takes a couple of minutes
    dt, _, _ = arm.run(sents[:min(len(sents), 32 * arm.cores)])
    rate = min(len(sents), 32 * arm.cores) / dt
    budget = max(1.0, min(12.0, 120.0 / (args.steps + args.warmup)))
    n_sample = int(max(arm.cores * 8, min(len(sents), rate * budget)))
    sample = sents[:n_sample]
    for _ in range(args.warmup):
        arm.run(sample)
    total = 0.0
    counters = {}
    for _ in range(args.steps):
        dt, _, counters = arm.run(sample)
        total += dt
    arm.close()
    value = len(sample) * args.steps / total
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args.config, cfg, len(sents)),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': arm.cores, 'kind': 'port',
                         'sample': '%d of %d sentences per step, multiprocessing.Pool(%d) over the pure-Python '
                                   'oracle port of Tagger.tag' % (len(sample), len(sents), arm.cores)},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'edges_per_sec': counters.get('E', 0) * args.steps / total if total else None,
        'transitions_per_sec': counters.get('T', 0) * args.steps / total if total else None,
    }
    emit(line)


def workload_config(name, cfg, n_sent):
    return {'workload': '%s: %d synthetic Hangul sentences (avg %d syllables), %d-morpheme synthetic dictionary, '
                        'beam=%d, %d trigram features' % (name, n_sent, cfg['mean_len'], cfg['n_dict'], cfg['beam'],
                                                           cfg['n_feat']),
            'sentences_per_gpu': n_sent, 'beam': cfg['beam'], 'l2': 'flushed between timed steps (256 MiB write)'}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    import lattice_based_tagger_b200 as pkg
    from lattice_based_tagger_b200 import _native, synth
    from lattice_based_tagger_b200.tagger.tagger import pack_sentences

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    cfg, dictionary, sents = synth.build_workload(args.config, rank=rank, n_sent=args.sentences)
    beam = cfg['beam'] if args.beam is None else args.beam
    reg = pkg.beam.BeamScoreFunctions(pkg.beam.RegularizationScore())
    reg_tagger = pkg.Tagger(dictionary, score_funcs=reg, device=local_rank)
    # the reference raises on sentences without any dictionary edge: replace them
    status = reg_tagger.tag_batch_packed(sents, beam)[3]
    good = [i for i in range(len(sents)) if status[i] == 0]
    for i in range(len(sents)):
        if status[i] != 0:
            sents[i] = sents[good[i % len(good)]]
    # features: rank 0's sample defines them for every rank (same tables everywhere)
    base_sents = sents if rank == 0 else synth.build_workload(args.config, rank=0, n_sent=args.sentences)[2]
    sample = feature_sample(base_sents)
    # (configurations whose feature target exceeds what lattice chains of the sample yield — C5 — are padded
    # with word n-grams over the dictionary)
    vocab = None
    if cfg['n_feat'] > 1_000_000:
        vocab = [(m, t) for t, ms in dictionary.tag_to_morphs.items() for m in sorted(ms)]
    feature_dic, coef = synth.make_features(
        sample, lambda s: reg_tagger.tag_batch(s, beam, errors='none'), reg_tagger.lattice_batch,
        cfg['n_feat'], list(dictionary.tag_to_morphs), seed=3, vocab=vocab)
    reg_tagger.close()
    funcs = pkg.beam.BeamScoreFunctions(
        pkg.beam.RegularizationScore(),
        pkg.beam.SimpleTrigramFeatureScore(pkg.features.SimpleTrigramEncoder(feature_dic), coef))
    tagger = pkg.Tagger(dictionary, score_funcs=funcs, device=local_rank)
    lib = tagger._lib
    batch = tagger._batch

    text, offsets = pack_sentences(sents)
    n = len(sents)
    n_units = int(offsets[-1])
    max_units = int(np.diff(offsets).max())
    d_text = torch.from_numpy(text.view(np.int16).copy()).to(dev)
    d_off = torch.from_numpy(offsets.copy()).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)
    sp = ctypes.c_void_p(stream.cuda_stream)

    p_text, p_off = ctypes.c_void_p(d_text.data_ptr()), ctypes.c_void_p(d_off.data_ptr())

    def step_device():
        _native.check(lib.lt_tag_batch_device(batch, p_text, p_off, n, n_units, max_units, beam, sp))

    # pinned host buffers for the end-to-end path
    def pinned(nbytes):
        return torch.empty(max(16, nbytes), dtype=torch.uint8).pin_memory()
    h_text = pinned(text.nbytes); h_text.numpy()[:text.nbytes] = text.view(np.uint8)
    h_off = pinned(offsets.nbytes); h_off.numpy()[:offsets.nbytes] = offsets.view(np.uint8)
    h_poff = pinned(4 * (n + 1)); h_edges = pinned(16 * max(1, n_units)); h_scores = pinned(8 * n); h_status = pinned(4 * n)

    def step_host():
        _native.check(lib.lt_tag_batch_host(
            batch, ctypes.c_void_p(h_text.data_ptr()), ctypes.c_void_p(h_off.data_ptr()), n, beam,
            ctypes.c_void_p(h_poff.data_ptr()), ctypes.c_void_p(h_edges.data_ptr()), n_units,
            ctypes.c_void_p(h_scores.data_ptr()), ctypes.c_void_p(h_status.data_ptr())))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    tagger.timings()                 # switches per-stage event timing on
    for _ in range(max(3, args.warmup)):
        step_device()
    torch.cuda.synchronize(dev)

    # ---- timed region: K steps, device-resident inputs ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    events = []
    stage = {'ms_lattice': 0.0, 'ms_beam': 0.0, 'ms_pack': 0.0}
    for _ in range(args.steps):
        flush.fill_(1)               # evict L2 between timed steps (outside the event bracket)
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step_device()
        e1.record(stream)
        events.append((e0, e1))
        e1.synchronize()
        t = tagger.timings()
        for k in stage:
            stage[k] += t[k]
    barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in events)
    counters = tagger.counters()

    # ---- end to end through the C ABI with host buffers ----
    for _ in range(2):
        step_host()
    barrier()
    host_s = 0.0
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        step_host()
        host_s += time.perf_counter() - t0
    barrier()
    clocks = sampler.stop()
    n_words = int(h_poff.numpy()[:4 * (n + 1)].view(np.int32)[n])

    ms_per_step = dev_ms / args.steps
    host_ms_per_step = 1e3 * host_s / args.steps
    rank_ms = [ms_per_step]
    if world > 1:
        mine = torch.tensor([ms_per_step, host_ms_per_step], dtype=torch.float64, device=dev)
        every = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(every, mine)
        rank_ms = [float(x[0]) for x in every]                 # reported per rank; the line's time is the maximum
        mhz = torch.tensor([float(clocks['sm_mhz'] or 0.0)], dtype=torch.float64, device=dev)
        all_mhz = [torch.zeros_like(mhz) for _ in range(world)]
        dist.all_gather(all_mhz, mhz)
        clocks['sm_mhz_by_rank'] = [float(x[0]) for x in all_mhz]
        ms_per_step, host_ms_per_step = max(rank_ms), max(float(x[1]) for x in every)
        c = torch.tensor([counters[k] for k in ('sentences', 'L', 'P', 'E', 'T', 'F', 'Bk', 'W')], dtype=torch.float64, device=dev)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        total_counters = dict(zip(('sentences', 'L', 'P', 'E', 'T', 'F', 'Bk', 'W'), [float(x) for x in c]))
    else:
        total_counters = {k: float(v) for k, v in counters.items()}

    if rank == 0:
        peak, peak_kind = measured_peak()
        c = counters
        bytes_lattice = 2 * c['L'] + 16 * c['P'] + 16 * c['E']
        bytes_beam = 16 * c['E'] + 16 * c['F'] + 8 * c['Bk'] + 8 * c['sentences'] + 4 * c['W']
        ms_lattice = stage['ms_lattice'] / args.steps
        ms_beam = stage['ms_beam'] / args.steps
        kernels = {
            'beam_kernel': {'achieved': bytes_beam / (ms_beam * 1e-3) / 1e9 if ms_beam else None,
                            'ms': ms_beam, 'algorithmic_bytes': bytes_beam},
            'lattice_kernel': {'achieved': bytes_lattice / (ms_lattice * 1e-3) / 1e9 if ms_lattice else None,
                                           'ms': ms_lattice, 'algorithmic_bytes': bytes_lattice},
        }
        dominant = 'beam_kernel' if ms_beam >= ms_lattice else 'lattice_kernel'
        traffic = None
        try:
            with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as f:
                traffic = json.load(f).get(dominant)
        except Exception:
            pass
        roofline = {'bound': 'hbm', 'kernel': dominant, 'achieved': kernels[dominant]['achieved'], 'peak': peak,
                    'peak_source': peak_kind, 'unit': 'GB/s',
                    'frac': (kernels[dominant]['achieved'] or 0.0) / peak, 'traffic': traffic, 'kernels': kernels}
        line = {
            'metric': METRIC, 'value': total_counters['sentences'] / (ms_per_step * 1e-3), 'unit': UNIT,
            'n_gpus': world, 'steps': args.steps, 'warmup': max(3, args.warmup), 'ms_per_step': ms_per_step,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': dict(workload_config(args.config, cfg, n), features=len(feature_dic)),
            'edges_per_sec': total_counters['E'] / (ms_per_step * 1e-3),
            'transitions_per_sec': total_counters['T'] / (ms_per_step * 1e-3),
            'e2e': {'value': total_counters['sentences'] / (host_ms_per_step * 1e-3), 'unit': UNIT,
                    'ms_per_step': host_ms_per_step,
                    'h2d_bytes_per_step': int(text.nbytes + offsets.nbytes),
                    'd2h_bytes_per_step': int(4 * (n + 1) + 16 * n_words + 8 * n + 4 * n)},
            # per step: batch prologue (zeroing + work order), lattice, beam, path-offset scan (one launch up to
            # 64 Ki sentences, else three), pack
            'gpu_launches': (5 if n + 1 <= 65536 else 7) * args.steps,
            'stage_ms_per_step': {k: v / args.steps for k, v in stage.items()},
            'ms_per_step_by_rank': rank_ms,
            'counters_per_step': counters,
            'roofline': roofline,
            'clocks': clocks,
            'tables_device_bytes': tagger._tables.device_bytes(),
        }
        if world == 1 and not args.no_cpu_baseline:
            line.update(cpu_baseline(args, cfg, sents, feature_dic, coef, beam, h_poff, h_edges, h_scores, h_status, n, tagger))
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(args, cfg, sents, feature_dic, coef, beam, h_poff, h_edges, h_scores, h_status, n, tagger):
    """Oracle port on this box's cores over a bounded sample, compared with the GPU results."""
    from lattice_based_tagger_b200 import _native
    arm = CpuArm(args.config, list(feature_dic.keys()), coef, beam)
    probe = sents[:min(len(sents), 16 * arm.cores)]
    dt, _, _ = arm.run(probe)
    rate = len(probe) / dt
    n_sample = int(max(len(probe), min(len(sents), rate * 12.0)))
    sample = sents[:n_sample]
    dt, results, totals = arm.run(sample)
    arm.close()
    # one process = one core (the reference itself is single-threaded)
    solo = CpuArm(args.config, list(feature_dic.keys()), coef, beam, cores=1)
    n_solo = int(max(32, min(len(sents), rate / arm.cores * 4.0)))
    solo_dt, _, _ = solo.run(sents[:n_solo])
    solo.close()
    # parity of the timed CPU sample against the GPU's last end-to-end step
    poff = h_poff.numpy()[:4 * (n + 1)].view(np.int32)
    edges = h_edges.numpy()[:16 * int(poff[n])].view(_native.EDGE_DTYPE)
    scores = h_scores.numpy()[:8 * n].view(np.float64)
    status = h_status.numpy()[:4 * n].view(np.int32)
    names = tagger._tables.tag_names
    mismatches = 0
    for i, want in enumerate(results):
        if want is None:
            mismatches += int(status[i] == 0)
            continue
        words = tagger.edges_to_words(sample[i].replace(' ', ''), edges[int(poff[i]):int(poff[i + 1])])
        got = [(w.b, w.e, w.tag0, w.morph0) for w in words]
        if got != want[0] or scores[i] != want[1] or status[i] != 0:
            mismatches += 1
    del names
    return {'cpu_baseline': {'value': len(sample) / dt, 'unit': UNIT, 'cores': arm.cores, 'kind': 'port',
                             'sample': '%d of %d sentences, multiprocessing.Pool(%d) over the pure-Python oracle '
                                       'port of Tagger.tag' % (len(sample), len(sents), arm.cores),
                             'transitions_per_sec': totals.get('T', 0) / dt,
                             'single_core': {'value': n_solo / solo_dt, 'unit': UNIT, 'sample': '%d sentences, 1 process' % n_solo}},
            'parity': {'checked': len(sample), 'mismatches': mismatches,
                       'what': 'segmentation, tags, lemmas and fp64 score bit-exact vs the CPU sample'}}


_RESULT_FD = None


def capture_stdout():
    """Everything libraries print on stdout (NCCL's version banner, for one) goes to stderr, so that
    stdout carries exactly the one JSON line."""
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + '\n').encode()
    sys.stdout.flush()
    if _RESULT_FD is None:
        os.write(1, data)
    else:
        os.write(_RESULT_FD, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', default='c2')
    ap.add_argument('--sentences', type=int, default=None, help='override the number of sentences per GPU')
    ap.add_argument('--beam', type=int, default=None)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    capture_stdout()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == '__main__':
    main()
