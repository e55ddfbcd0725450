#!/usr/bin/env python
"""Benchmark of the batched decode path (lattice build + feature scoring + beam search).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c2] [--impl reference]

One step = one pass of the hot path over one batch of synthetic sentences (SURVEY.md §8d
generators).  At N=1 the workload is BASELINE.json configs[1] (C2): 10k synthetic Hangul sentences
(avg 20 syllables), 100k-morpheme synthetic dictionary, beam=5.

Printed JSON line (rank 0):
  value         sentences/s with inputs resident in HBM (CUDA events around the two stage launches; the
                result fetch is not inside — `e2e` has it); under torchrun every rank tags its own batch
                (weak scaling, no collective on the data path) and the step time is the max over ranks
  e2e           the same through `lt_tag_batch_host` with pinned host buffers, copies inside the timed region
  roofline      algorithmic bytes (device work counters) of the dominant kernel / its device time / the
                measured HBM peak; `kernels` has both kernels
  cpu_baseline  the reference's own Python `Tagger.tag` (vendored unmodified to oracle/_ref by
                `oracle/vendor_ref.py`, else the oracle's port of it) on this box's cores, with a
                bit-exact comparison of every timed sentence against the GPU result
  api           the Python API a drop-in user calls: `Tagger.tag_batch` (lazy and fully materialised) and
                single-sentence `Tagger.tag` latency, C1 = the bundled `base` dictionary, beam 5
  other_configs (N=1) bounded samples of BASELINE configs C3 / C4 / C5 in the same process: step and
                stage times, per-kernel roofline, parity against the CPU arm
  strong        (N>1) ONE corpus partitioned by estimated work over the ranks (`sharding.py`), tagged
                through `lt_tag_batch_host`, packed results gathered on rank 0 in input order and compared
                with a single-GPU pass
`--impl reference` times the CPU arm alone, with all host cores.
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
COUNTER_NAMES = ('sentences', 'L', 'P', 'E', 'T', 'F', 'Bk', 'W')

# bounded samples of the larger configurations (full dictionary / feature-table sizes, fewer sentences)
OTHER_SAMPLES = {'c3': 20_000, 'c4': 1_000, 'c5': 20_000}


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


def measured_traffic(config):
    """DRAM bytes per launch of the two kernels from an `ncu --set full` capture of this
    configuration (profiles/traffic.json, written by profiles/tools/ncu_summary.py), or None."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as f:
            return json.load(f).get(config)
    except Exception:
        return None


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


def workload_config(name, cfg, n_sent):
    return {'workload': '%s: %d synthetic Hangul sentences (avg %d syllables), %d-morpheme synthetic dictionary, '
                        'beam=%d, %d trigram features' % (name, n_sent, cfg['mean_len'], cfg['n_dict'], cfg['beam'],
                                                           cfg['n_feat']),
            'sentences_per_gpu': n_sent, 'beam': cfg['beam'], 'features': cfg['n_feat'],
            'l2': 'flushed between timed steps (256 MiB write)'}


def k3_first_flags(rules):
    """Iteration order of {2-syllable key, 3-syllable key} as this process's string hashing gives it
    (lemmatizer.py:107) — handed to every tagger / oracle of the run so that they all agree."""
    flags = {}
    for key in rules:
        if len(key) == 3 and key[:2] in rules:
            flags[key] = next(iter({key[:2], key})) == key
    return flags


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's Python path (vendored copy, else the oracle's port) — the only place
# bench.py touches oracle/
# ------------------------------------------------------------------------------------------------
_W = {}


def cpu_kind():
    from oracle import vendor_ref
    return 'reference' if vendor_ref.vendor() else 'port'


def _synthetic_dictionary(config):
    from lattice_based_tagger_b200 import synth
    cfg = dict(synth.CONFIGS[config])
    alphabet = synth.make_alphabet(cfg['alphabet'], seed=0)
    tag_to_morphs = synth.make_dictionary(cfg['n_dict'], alphabet, seed=0)
    rules = synth.make_rules(tag_to_morphs, alphabet, n_keys=min(5000, max(50, cfg['n_dict'] // 20)), seed=1)
    return tag_to_morphs, rules


def _cpu_tagger(kind, tag_to_morphs, rules, feature_items, coef, reg_params=None):
    """-> tag(sent, beam) returning ([(b, e, tag0, morph0)], score) or None for the reference's IndexError."""
    reg_params = reg_params or {}
    if kind == 'reference':
        from oracle import vendor_ref
        ref = vendor_ref.import_reference()
        dictionary = ref.dictionary.MorphemeDictionary(tag_to_morphs, rules)
        funcs = [ref.beam.RegularizationScore(**reg_params)]
        if feature_items is not None:
            feature_dic = {tuple(k): i for i, k in enumerate(feature_items)}
            funcs.append(ref.beam.SimpleTrigramFeatureScore(ref.features.SimpleTrigramEncoder(feature_dic), coef))
        tagger = ref.tagger.Tagger(dictionary, score_funcs=ref.beam.BeamScoreFunctions(*funcs))

        def tag(sent, beam):
            try:
                seq = tagger.tag(sent, beam_size=beam)
            except IndexError:
                return None
            return [(w.b, w.e, w.tag0, w.morph0) for w in seq.sequences[1:-1]], seq.score
        return tag, None
    import lattice_based_tagger_b200 as pkg
    from oracle import lattice_oracle as lo
    dictionary = pkg.dictionary.MorphemeDictionary(tag_to_morphs, rules)
    funcs = [pkg.beam.RegularizationScore(**reg_params)]
    if feature_items is not None:
        feature_dic = {tuple(k): i for i, k in enumerate(feature_items)}
        funcs.append(pkg.beam.SimpleTrigramFeatureScore(pkg.features.SimpleTrigramEncoder(feature_dic), coef))
    tagger = lo.OracleTagger(dictionary, pkg.beam.BeamScoreFunctions(*funcs))

    def tag(sent, beam):
        try:
            best = tagger.tag(sent, beam)
        except IndexError:
            return None
        return [(w[6], w[7], w[3], w[1]) for w in best.words[1:-1]], best.score
    return tag, tagger


def _worker_init(kind, config, dict_payload, feature_items, coef_bytes, beam):
    if dict_payload is None:
        tag_to_morphs, rules = _synthetic_dictionary(config)
    else:
        tag_to_morphs = {t: set(m) for t, m in dict_payload[0]}
        rules = {k: tuple(tuple(c) for c in v) for k, v in dict_payload[1]}
    coef = None if coef_bytes is None else np.frombuffer(coef_bytes, dtype=np.float64)
    _W['tag'], _ = _cpu_tagger(kind, tag_to_morphs, rules, feature_items, coef)
    _W['beam'] = beam


def _worker_tag(sents):
    tag, beam = _W['tag'], _W['beam']
    return [tag(sent, beam) for sent in sents]


class CpuArm:
    """multiprocessing.Pool over the host cores running the reference's Python `Tagger.tag`."""

    def __init__(self, config, feature_items, coef, beam, cores=None, kind=None, dict_payload=None):
        import multiprocessing as mp
        self.kind = kind or cpu_kind()
        self.cores = cores or os.cpu_count() or 1
        ctx = mp.get_context('spawn')        # (PYTHONHASHSEED=0 is inherited: same set orders in every process)
        coef_bytes = None if coef is None else np.asarray(coef, dtype=np.float64).tobytes()
        self.pool = ctx.Pool(self.cores, initializer=_worker_init,
                             initargs=(self.kind, config, dict_payload, feature_items, coef_bytes, beam))
        # make sure every worker finished building its tables before anything is timed
        self.pool.map(_worker_tag, [[] for _ in range(self.cores * 2)])

    def describe(self):
        if self.kind == 'reference':
            return "the reference's own Tagger.tag (unmodified copy in oracle/_ref)"
        return "the pure-Python oracle port of Tagger.tag"

    def run(self, sents):
        """-> (seconds, results) for tagging `sents` across the pool."""
        chunk = max(1, min(64, len(sents) // (self.cores * 4) or 1))
        chunks = [sents[i:i + chunk] for i in range(0, len(sents), chunk)]
        t0 = time.perf_counter()
        parts = self.pool.map(_worker_tag, chunks)
        dt = time.perf_counter() - t0
        return dt, [r for part in parts for r in part]

    def close(self):
        self.pool.close()
        self.pool.join()


def feature_sample(sents):
    return sents[:max(64, min(2000, len(sents) // 10))]


# ------------------------------------------------------------------------------------------------
# workloads (GPU side)
# ------------------------------------------------------------------------------------------------
class Workload:
    pass


def prepare_workload(config, rank, n_sent, device, beam_override=None):
    """Dictionary, sentences, features and the tagger of one configuration.  Sentences are those of
    `rank`; the feature dictionary always comes from rank 0's sentences, so every rank holds the same
    tables."""
    import lattice_based_tagger_b200 as pkg
    from lattice_based_tagger_b200 import synth
    W = Workload()
    W.config = config
    W.cfg, W.dictionary, sents0 = synth.build_workload(config, rank=0, n_sent=n_sent)
    W.beam = W.cfg['beam'] if beam_override is None else beam_override
    W.k3 = k3_first_flags(W.dictionary.rules)
    reg = pkg.beam.BeamScoreFunctions(pkg.beam.RegularizationScore())
    reg_tagger = pkg.Tagger(W.dictionary, score_funcs=reg, device=device, k3_first=W.k3)

    def replace_unhit(sents):
        # the reference raises on sentences without any dictionary edge: replace them
        status = reg_tagger.tag_batch_packed(sents, W.beam)[3]
        good = [i for i in range(len(sents)) if status[i] == 0]
        for i in range(len(sents)):
            if status[i] != 0:
                sents[i] = sents[good[i % len(good)]]
        return sents

    sents0 = replace_unhit(sents0)
    sample = feature_sample(sents0)
    # (configurations whose feature target exceeds what lattice chains of the sample yield — C5 — are padded
    # with word n-grams over the dictionary)
    vocab = None
    if W.cfg['n_feat'] > 1_000_000:
        vocab = [(m, t) for t, ms in W.dictionary.tag_to_morphs.items() for m in sorted(ms)]
    W.feature_dic, W.coef = synth.make_features(
        sample, lambda s: reg_tagger.tag_batch(s, W.beam, errors='none'), reg_tagger.lattice_batch,
        W.cfg['n_feat'], list(W.dictionary.tag_to_morphs), seed=3, vocab=vocab)
    if rank == 0:
        W.sents = sents0
    else:
        W.sents = replace_unhit(synth.build_workload(config, rank=rank, n_sent=n_sent)[2])
    W.reg_tagger = reg_tagger
    funcs = pkg.beam.BeamScoreFunctions(
        pkg.beam.RegularizationScore(),
        pkg.beam.SimpleTrigramFeatureScore(pkg.features.SimpleTrigramEncoder(W.feature_dic), W.coef))
    W.funcs = funcs
    W.tagger = pkg.Tagger(W.dictionary, score_funcs=funcs, device=device, k3_first=W.k3)
    return W


class DeviceBatch:
    """One batch's buffers: the packed text resident in HBM, pinned host buffers for the end-to-end path."""

    def __init__(self, sents, dev):
        import torch
        from lattice_based_tagger_b200.engine import pack_sentences
        self.sents = sents
        self.text, self.offsets = pack_sentences(sents)
        self.n = len(sents)
        self.n_units = int(self.offsets[-1])
        self.max_units = int(np.diff(self.offsets).max()) if self.n else 0
        self.d_text = torch.from_numpy(self.text.view(np.int16).copy()).to(dev)
        self.d_off = torch.from_numpy(self.offsets.copy()).to(dev)

        def pinned(nbytes):
            return torch.empty(max(16, nbytes), dtype=torch.uint8).pin_memory()
        self.h_text = pinned(self.text.nbytes)
        self.h_text.numpy()[:self.text.nbytes] = self.text.view(np.uint8)
        self.h_off = pinned(self.offsets.nbytes)
        self.h_off.numpy()[:self.offsets.nbytes] = self.offsets.view(np.uint8)
        self.h_poff = pinned(4 * (self.n + 1))
        self.h_edges = pinned(16 * max(1, self.n_units))
        self.h_scores = pinned(8 * self.n)
        self.h_status = pinned(4 * self.n)

    def results(self):
        from lattice_based_tagger_b200 import _native
        n = self.n
        poff = self.h_poff.numpy()[:4 * (n + 1)].view(np.int32)
        edges = self.h_edges.numpy()[:16 * int(poff[n])].view(_native.EDGE_DTYPE)
        scores = self.h_scores.numpy()[:8 * n].view(np.float64)
        status = self.h_status.numpy()[:4 * n].view(np.int32)
        return poff, edges, scores, status


def time_workload(W, B, steps, warmup, dev, barrier, flush):
    """Device-resident and end-to-end timing of one workload on this rank.
    -> dict(ms_per_step, host_ms_per_step, stage ms per step, counters, launches per step, info)."""
    import torch
    from lattice_based_tagger_b200 import _native
    tagger, lib, batch = W.tagger, W.tagger._lib, W.tagger._batch
    stream = torch.cuda.current_stream(dev)
    sp = ctypes.c_void_p(stream.cuda_stream)
    p_text, p_off = ctypes.c_void_p(B.d_text.data_ptr()), ctypes.c_void_p(B.d_off.data_ptr())

    def step_device():
        _native.check(lib.lt_tag_batch_device(batch, p_text, p_off, B.n, B.n_units, B.max_units, W.beam, sp))

    def step_host():
        _native.check(lib.lt_tag_batch_host(
            batch, ctypes.c_void_p(B.h_text.data_ptr()), ctypes.c_void_p(B.h_off.data_ptr()), B.n, W.beam,
            ctypes.c_void_p(B.h_poff.data_ptr()), ctypes.c_void_p(B.h_edges.data_ptr()), max(1, B.n_units),
            ctypes.c_void_p(B.h_scores.data_ptr()), ctypes.c_void_p(B.h_status.data_ptr())))

    # warm-up: every step is RESOLVED (overflow flags read back, buffers grown, stages rerun, staging capacity
    # adapted), so that all capacities have settled before anything is timed
    tagger.set_stage_timing(False)
    for _ in range(max(3, warmup)):
        step_device()
        tagger.info()
    step_device()
    torch.cuda.synchronize(dev)

    def timed_pass(staged):
        """`steps` timed steps; staged = per-stage events between the kernels of a step (which then run strictly
        one after the other), else the kernels are programmatic dependent launches."""
        tagger.set_stage_timing(staged)
        step_device()
        torch.cuda.synchronize(dev)
        for attempt in range(2):
            info0 = tagger.info()
            barrier()
            events = []
            stage = {'ms_lattice': 0.0, 'ms_beam': 0.0, 'ms_pack': 0.0}
            for _ in range(steps):
                flush.fill_(1)               # evict L2 between timed steps (outside the event bracket)
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                step_device()
                e1.record(stream)
                e1.synchronize()
                bracket = e0.elapsed_time(e1)
                if staged:
                    t = tagger.timings()
                    inside = t['ms_lattice'] + t['ms_beam'] + t['ms_pack']
                    # the stages are timed by events inside the bracket: their sum cannot exceed it
                    if inside > bracket * 1.02 + 0.02:
                        raise RuntimeError('stage times %.4f ms exceed the step bracket %.4f ms: part of the step ran '
                                           'outside the timed region' % (inside, bracket))
                    for k in stage:
                        stage[k] += t[k]
                events.append(bracket)
            barrier()
            info1 = tagger.info()
            if info1['reruns'] == info0['reruns']:
                break
            if attempt == 1:
                raise RuntimeError('lattice buffers were still growing inside the timed region (%d reruns)'
                                   % (info1['reruns'] - info0['reruns']))
        return events, stage, info0, info1

    out = {}
    # headline: the step as a user runs it (no events between its kernels)
    events, _, info0, info1 = timed_pass(False)
    out['ms_per_step'] = sum(events) / steps
    out['launches_per_step'] = (info1['launches'] - info0['launches']) / steps
    out['reruns_in_timed_region'] = info1['reruns'] - info0['reruns']
    # per-kernel durations for the roofline: the same steps with CUDA events between the kernels
    events_s, stage, s0, s1 = timed_pass(True)
    out['ms_per_step_staged'] = sum(events_s) / steps
    out['stage'] = {k: v / steps for k, v in stage.items()}
    out['reruns_in_timed_region'] += s1['reruns'] - s0['reruns']
    out['counters'] = tagger.counters()
    out['info'] = s1
    tagger.set_stage_timing(False)

    # ---- end to end through the C ABI with host buffers ----
    for _ in range(2):
        step_host()
    barrier()
    host_s = 0.0
    for _ in range(steps):
        flush.fill_(1)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        step_host()
        host_s += time.perf_counter() - t0
    barrier()
    out['host_ms_per_step'] = 1e3 * host_s / steps
    out['n_words'] = int(B.h_poff.numpy()[:4 * (B.n + 1)].view(np.int32)[B.n])
    # digest of everything the last end-to-end step returned: equal digests = bit-identical results (used to
    # compare kernel variants with a build that was checked against the CPU arm)
    import hashlib
    digest = hashlib.sha1()
    for part in B.results():
        digest.update(np.ascontiguousarray(part).view(np.uint8).tobytes())
    out['results_sha1'] = digest.hexdigest()
    return out


def roofline_of(config, counters, stage, peak, peak_kind):
    c = counters
    bytes_lattice = 2 * c['L'] + 16 * c['P'] + 16 * c['E']
    bytes_beam = 16 * c['E'] + 16 * c['F'] + 8 * c['Bk'] + 8 * c['sentences'] + 4 * c['W']
    ms_lattice, ms_beam = stage['ms_lattice'], stage['ms_beam']
    traffic = measured_traffic(config) or {}
    kernels = {
        'beam_kernel': {'achieved': bytes_beam / (ms_beam * 1e-3) / 1e9 if ms_beam else None, 'ms': ms_beam,
                        'algorithmic_bytes': bytes_beam, 'traffic': traffic.get('beam_kernel')},
        'lattice_kernel': {'achieved': bytes_lattice / (ms_lattice * 1e-3) / 1e9 if ms_lattice else None, 'ms': ms_lattice,
                           'algorithmic_bytes': bytes_lattice, 'traffic': traffic.get('lattice_kernel')},
    }
    for k in kernels.values():
        k['frac'] = (k['achieved'] or 0.0) / peak
    dominant = 'beam_kernel' if ms_beam >= ms_lattice else 'lattice_kernel'
    return {'bound': 'hbm', 'kernel': dominant, 'achieved': kernels[dominant]['achieved'], 'peak': peak,
            'peak_source': peak_kind, 'unit': 'GB/s', 'frac': kernels[dominant]['frac'],
            'traffic': kernels[dominant]['traffic'], 'kernels': kernels}


def parity_against_cpu(W, B, results, sample):
    """Bit-exact comparison of the CPU arm's results with the GPU's last end-to-end step."""
    poff, edges, scores, status = B.results()
    mismatches = 0
    for i, want in enumerate(results):
        if want is None:
            mismatches += int(status[i] == 0)
            continue
        words = W.tagger.edges_to_words(sample[i].replace(' ', ''), edges[int(poff[i]):int(poff[i + 1])])
        got = [(w.b, w.e, w.tag0, w.morph0) for w in words]
        if got != want[0] or scores[i] != want[1] or status[i] != 0:
            mismatches += 1
    return {'checked': len(results), 'mismatches': mismatches,
            'what': 'segmentation, tags, lemmas and fp64 score bit-exact vs the CPU sample'}


def cpu_baseline(W, B, seconds=12.0, with_single_core=True):
    """The CPU arm on this box's cores over a bounded sample of the same workload, compared with the GPU results."""
    items = list(W.feature_dic.keys())
    arm = CpuArm(W.config, items, W.coef, W.beam)
    probe = W.sents[:min(len(W.sents), 8 * arm.cores)]
    dt, _ = arm.run(probe)
    rate = len(probe) / dt
    n_sample = int(max(len(probe), min(len(W.sents), rate * seconds)))
    sample = W.sents[:n_sample]
    dt, results = arm.run(sample)
    arm.close()
    out = {'cpu_baseline': {'value': len(sample) / dt, 'unit': UNIT, 'cores': arm.cores, 'kind': arm.kind,
                            'sample': '%d of %d sentences, multiprocessing.Pool(%d) over %s'
                                      % (len(sample), len(W.sents), arm.cores, arm.describe())},
           'parity': parity_against_cpu(W, B, results, sample)}
    if with_single_core:
        # one process = one core (the reference itself is single-threaded)
        solo = CpuArm(W.config, items, W.coef, W.beam, cores=1, kind=arm.kind)
        n_solo = int(max(32, min(len(W.sents), rate / arm.cores * 4.0)))
        solo_dt, _ = solo.run(W.sents[:n_solo])
        solo.close()
        out['cpu_baseline']['single_core'] = {'value': n_solo / solo_dt, 'unit': UNIT, 'sample': '%d sentences, 1 process' % n_solo}
    return out


# ------------------------------------------------------------------------------------------------
# the Python API a drop-in user calls
# ------------------------------------------------------------------------------------------------
def api_numbers(W, device):
    import lattice_based_tagger_b200 as pkg
    out = {}
    tagger = W.tagger
    sents = W.sents
    for _ in range(2):
        tagger.tag_batch(sents, W.beam)
    reps = 5
    t0 = time.perf_counter()
    for _ in range(reps):
        res = tagger.tag_batch(sents, W.beam)
    lazy = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    for _ in range(2):
        full = tagger.tag_batch(sents, W.beam, lazy=False)
    eager = (time.perf_counter() - t0) / 2
    assert len(res) == len(full) == len(sents)
    out['tag_batch_sent_per_s'] = len(sents) / lazy
    out['tag_batch_materialized_sent_per_s'] = len(sents) / eager
    out['tag_batch_what'] = ('Tagger.tag_batch over the %d %s sentences from Python strings: packing + lt_tag_batch_host + a '
                             'PackedSequences result (Word tuples built on access); "materialized" builds every Sequence / Word'
                             % (len(sents), W.config))
    one = sents[0]
    for _ in range(20):
        tagger.tag(one, W.beam)
    t0 = time.perf_counter()
    for i in range(200):
        tagger.tag(sents[i % len(sents)], W.beam)
    out['tag_single_ms'] = 1e3 * (time.perf_counter() - t0) / 200
    out['c1'] = c1_numbers(pkg, device)
    return out


def c1_numbers(pkg, device):
    """BASELINE configs[0]: single-sentence `Tagger.tag` with the bundled `base` dictionary, beam 5 — this
    package on the GPU next to the reference's Python on one core, same sentences."""
    from tests import _cases, _golden
    payload = _golden.load('base_c1')            # the `base` dictionary and feature weights as the reference loaded them
    case = payload['case']
    dictionary, funcs = _cases.build_objects(case, pkg)
    tagger = pkg.Tagger(dictionary, score_funcs=funcs, device=device, k3_first=payload['k3_first'])
    sents = [s for s in case['sentences'] if s]
    for s in sents:
        tagger.tag(s, 5)
    t0 = time.perf_counter()
    reps = 20
    for _ in range(reps):
        for s in sents:
            tagger.tag(s, 5)
    gpu_ms = 1e3 * (time.perf_counter() - t0) / (reps * len(sents))
    t0 = time.perf_counter()
    batch = tagger.tag_batch(sents * 100, 5, lazy=False)
    batch_ms = 1e3 * (time.perf_counter() - t0) / len(batch)
    out = {'workload': 'c1: %d Korean sentences, bundled base dictionary (%d entries), beam=5'
                       % (len(sents), sum(len(m) for m in dictionary.tag_to_morphs.values())),
           'tag_single_ms': gpu_ms, 'tag_batch_ms_per_sentence': batch_ms}
    kind = cpu_kind()
    tag_to_morphs = {t: set(case['tag_to_morphs'][t]) for t in case['tags']}
    rules = {k: tuple(tuple(c) for c in v) for k, v in case['rules'].items()}
    # (the same scorers as the fixture: RegularizationScore + trigram features)
    reg = next(f for f in case['funcs'] if f['kind'] == 'reg')
    params = {k: reg[k] for k in ('unknown_penalty', 'known_preference', 'syllable_penalty')}
    tag, _ = _cpu_tagger(kind, tag_to_morphs, rules, case['feature_keys'], np.asarray(case['coefficients'], dtype=np.float64), params)
    for s in sents[:3]:
        tag(s, 5)
    t0 = time.perf_counter()
    ref_results = [tag(s, 5) for s in sents]
    ref_ms = 1e3 * (time.perf_counter() - t0) / len(sents)
    same = 0
    for s, want in zip(sents, ref_results):
        got = tagger.tag(s, 5)
        same += int(want is not None and [(w.b, w.e, w.tag0, w.morph0) for w in got.sequences[1:-1]] == want[0] and got.score == want[1])
    out['reference'] = {'tag_single_ms': ref_ms, 'kind': kind, 'cores': 1}
    out['parity'] = {'checked': len(sents), 'identical': same}
    tagger.close()
    return out


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    peak, peak_kind = measured_peak()

    W = prepare_workload(args.config, rank, args.sentences, local_rank, args.beam)
    B = DeviceBatch(W.sents, dev)
    sampler = ClockSampler(local_rank)
    sampler.start()
    R = time_workload(W, B, args.steps, args.warmup, dev, barrier, flush)
    clocks = sampler.stop()

    ms_per_step, host_ms_per_step = R['ms_per_step'], R['host_ms_per_step']
    counters = R['counters']
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
        c = torch.tensor([counters[k] for k in COUNTER_NAMES], dtype=torch.float64, device=dev)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        total_counters = dict(zip(COUNTER_NAMES, [float(x) for x in c]))
    else:
        total_counters = {k: float(v) for k, v in counters.items()}

    strong = None
    if world > 1 and not args.no_strong:
        strong = strong_scaling(args, W, dev, rank, world, local_rank, flush)

    if rank == 0:
        n = B.n
        line = {
            'metric': METRIC, 'value': total_counters['sentences'] / (ms_per_step * 1e-3), 'unit': UNIT,
            'n_gpus': world, 'steps': args.steps, 'warmup': max(3, args.warmup) + 1, 'ms_per_step': ms_per_step,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': workload_config(args.config, W.cfg, n),
            'value_excludes': 'result fetch (device-to-host copy of the paths); e2e includes it',
            'edges_per_sec': total_counters['E'] / (ms_per_step * 1e-3),
            'transitions_per_sec': total_counters['T'] / (ms_per_step * 1e-3),
            'e2e': {'value': total_counters['sentences'] / (host_ms_per_step * 1e-3), 'unit': UNIT,
                    'ms_per_step': host_ms_per_step,
                    'h2d_bytes_per_step': int(B.text.nbytes + B.offsets.nbytes),
                    'd2h_bytes_per_step': int(4 * (n + 1) + 16 * R['n_words'] + 8 * n + 4 * n)},
            # counted by the library (lt_batch_info): batch prologue (zeroing + work order), lattice, beam,
            # path-offset scan (one launch up to 64 Ki sentences, else three), pack
            'gpu_launches': int(round(R['launches_per_step'] * args.steps)),
            # per-kernel durations by CUDA events BETWEEN the kernels of a step, taken over a second set of timed steps:
            # with those events the kernels run strictly one after the other (`ms_per_step_staged` is the step in that
            # mode); the headline step has none, so each kernel's launch overlaps its predecessor's tail (PDL)
            'stage_ms_per_step': R['stage'],
            'ms_per_step_staged': R['ms_per_step_staged'],
            'reruns_in_timed_region': R['reruns_in_timed_region'],
            'results_sha1': R['results_sha1'],
            'ms_per_step_by_rank': rank_ms,
            'counters_per_step': counters,
            'roofline': roofline_of(args.config, counters, R['stage'], peak, peak_kind),
            'launch': {k: R['info'][k] for k in ('hcap', 'retry_hcap', 'retried', 'edge_cap', 'lattice_warps', 'lattice_ctas_per_sm',
                                                  'lattice_smem', 'beam_warps', 'beam_ctas_per_sm', 'beam_smem', 'beam_trail_smem')},
            'clocks': clocks,
            'tables_device_bytes': W.tagger._tables.device_bytes(),
        }
        if strong is not None:
            line['strong'] = strong
        if world == 1 and not args.no_cpu_baseline:
            line.update(cpu_baseline(W, B))
        # (the sections after the headline are additions to it: a failure in one of them is reported in its place
        # and must not cost the line)
        if world == 1 and not args.no_api:
            try:
                line['api'] = api_numbers(W, local_rank)
            except Exception as exc:
                line['api'] = {'error': '%s: %s' % (type(exc).__name__, exc)}
        if world == 1 and args.other_configs:
            W.tagger.close()
            W.reg_tagger.close()
            del B
            line['other_configs'] = {}
            for name in args.other_configs:
                try:
                    line['other_configs'][name] = other_config(name, args, dev, local_rank, barrier, flush, peak, peak_kind)
                except Exception as exc:
                    line['other_configs'][name] = {'error': '%s: %s' % (type(exc).__name__, exc)}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def other_config(name, args, dev, local_rank, barrier, flush, peak, peak_kind):
    """A bounded sample of one of the larger configurations: full dictionary / feature-table sizes, fewer
    sentences.  Same timing rules as the headline (resolved warm-up, stage sum <= bracket, no rerun inside)."""
    t0 = time.perf_counter()
    n_sent = OTHER_SAMPLES.get(name, 10_000)
    W = prepare_workload(name, 0, n_sent, local_rank)
    B = DeviceBatch(W.sents, dev)
    build_s = time.perf_counter() - t0
    steps = max(3, min(args.steps, 5))
    R = time_workload(W, B, steps, 3, dev, barrier, flush)
    c = R['counters']
    out = {
        'config': workload_config(name, W.cfg, B.n), 'steps': steps,
        'ms_per_step': R['ms_per_step'], 'value': B.n / (R['ms_per_step'] * 1e-3), 'unit': UNIT,
        'edges_per_sec': c['E'] / (R['ms_per_step'] * 1e-3), 'transitions_per_sec': c['T'] / (R['ms_per_step'] * 1e-3),
        'e2e': {'value': B.n / (R['host_ms_per_step'] * 1e-3), 'unit': UNIT, 'ms_per_step': R['host_ms_per_step']},
        'stage_ms_per_step': R['stage'], 'ms_per_step_staged': R['ms_per_step_staged'],
        'reruns_in_timed_region': R['reruns_in_timed_region'],
        'results_sha1': R['results_sha1'], 'counters_per_step': c, 'roofline': roofline_of(name, c, R['stage'], peak, peak_kind),
        'launch': {k: R['info'][k] for k in ('hcap', 'retry_hcap', 'retried', 'lattice_warps', 'lattice_ctas_per_sm', 'beam_warps',
                                              'beam_ctas_per_sm', 'beam_trail_smem')},
        'tables_device_bytes': W.tagger._tables.device_bytes(), 'build_seconds': build_s,
    }
    if not args.no_cpu_baseline:
        out.update(cpu_baseline(W, B, seconds=5.0, with_single_core=False))
    if name == 'c5':
        # the beam-width sweep of BASELINE configs[4] on the same tables (1 = pure Viterbi .. 64)
        sweep = {}
        for k in (1, 8, 64):
            W.beam = k
            Rk = time_workload(W, B, 3, 3, dev, barrier, flush)
            sweep[str(k)] = {'ms_per_step': Rk['ms_per_step'], 'value': B.n / (Rk['ms_per_step'] * 1e-3),
                             'stage_ms_per_step': Rk['stage'],
                             'beam_frac': roofline_of(name, Rk['counters'], Rk['stage'], peak, peak_kind)['kernels']['beam_kernel']['frac']}
        out['beam_sweep'] = sweep
    W.tagger.close()
    W.reg_tagger.close()
    return out


# ------------------------------------------------------------------------------------------------
# strong scaling: one corpus, sharded
# ------------------------------------------------------------------------------------------------
_CORPUS = {}


def _corpus_chunk(job):
    from lattice_based_tagger_b200 import synth
    config, chunk, n = job
    if config not in _CORPUS:            # one dictionary per worker process, not per chunk
        cfg = dict(synth.CONFIGS[config])
        alphabet = synth.make_alphabet(cfg['alphabet'], seed=0)
        tag_to_morphs = synth.make_dictionary(cfg['n_dict'], alphabet, seed=0)
        rules = synth.make_rules(tag_to_morphs, alphabet, n_keys=min(5000, max(50, cfg['n_dict'] // 20)), seed=1)
        _CORPUS[config] = (cfg, alphabet, tag_to_morphs, rules)
    cfg, alphabet, tag_to_morphs, rules = _CORPUS[config]
    return synth.make_sentences(tag_to_morphs, rules, n, cfg['mean_len'], alphabet, seed=7000 + chunk, fixed_len=cfg['fixed_len'])


def strong_scaling(args, W_headline, dev, rank, world, local_rank, flush):
    """ONE corpus (rank 0 builds it, the packed text is broadcast), partitioned by estimated work, every rank
    tags its shard through `lt_tag_batch_host`, the packed results are gathered on rank 0 in input order
    (numpy arrays, no Python objects) and compared with rank 0's own single-GPU pass over the whole corpus."""
    import torch
    import torch.distributed as dist
    from lattice_based_tagger_b200 import _native, sharding
    n_total = args.strong_sentences
    config = args.strong_config
    # the tables of the sharded configuration on every rank (BASELINE configs[2]: C3); its feature dictionary
    # comes from a 20k-sentence sample of the generator, identical on every rank
    W = W_headline if config == args.config else prepare_workload(config, 0, OTHER_SAMPLES.get(config, 10_000), local_rank)
    t_build = time.perf_counter()
    if rank == 0:
        import multiprocessing as mp
        chunk = 10_000
        jobs = [(config, c, min(chunk, n_total - c * chunk)) for c in range((n_total + chunk - 1) // chunk)]
        with mp.get_context('spawn').Pool(min(len(jobs), max(1, (os.cpu_count() or 2) // 2))) as pool:
            corpus = [s for part in pool.map(_corpus_chunk, jobs) for s in part]
        # sentences no dictionary word covers raise in the reference: keep them, they carry status 1 through the gather
        from lattice_based_tagger_b200.engine import pack_sentences
        text, offsets = pack_sentences(corpus)
        header = torch.tensor([text.size, offsets.size], dtype=torch.int64, device=dev)
    else:
        header = torch.zeros(2, dtype=torch.int64, device=dev)
    dist.broadcast(header, 0)
    n_text, n_off = int(header[0]), int(header[1])
    if rank == 0:
        d_text = torch.from_numpy(text.view(np.int16).copy()).to(dev)
        d_off = torch.from_numpy(offsets.copy()).to(dev)
    else:
        d_text = torch.empty(n_text, dtype=torch.int16, device=dev)
        d_off = torch.empty(n_off, dtype=torch.int32, device=dev)
    dist.broadcast(d_text.view(torch.uint8), 0)          # (NCCL has no 16-bit integer type: the code units travel as bytes)
    dist.broadcast(d_off, 0)
    text = d_text.cpu().numpy().view(np.uint16)
    offsets = d_off.cpu().numpy()
    build_s = time.perf_counter() - t_build
    n = offsets.size - 1
    lengths = np.diff(offsets).astype(np.int64)

    # partition by estimated work (transitions ~ length x (8 k + k E / L), SURVEY §8e: the batch's beam and lattice
    # density are common factors): CONTIGUOUS slices of equal total length, so that the gathered results are a plain
    # concatenation in input order (sharding.partition_by_work deals sorted sentences instead; it balances the
    # length mix too but needs an inverse permutation of every word record at the gather)
    bounds = sharding.shard_bounds(lengths, world)
    parts = [np.arange(bounds[r], bounds[r + 1], dtype=np.int64) for r in range(world)]
    mine = parts[rank]
    lib, batch = W.tagger._lib, W.tagger._batch

    def tag_indices(idx):
        """-> (path_len per sentence, path edge records, scores, status) of sentences `idx`, in that order."""
        sub_off = np.zeros(idx.size + 1, dtype=np.int64)
        np.cumsum(lengths[idx], out=sub_off[1:])
        sub_text = np.empty(max(1, int(sub_off[-1])), dtype=np.uint16)
        starts = offsets[idx].astype(np.int64)
        gather = np.repeat(starts - sub_off[:-1], lengths[idx]) + np.arange(int(sub_off[-1]), dtype=np.int64)
        sub_text[:gather.size] = text[gather]
        sub_off32 = sub_off.astype(np.int32)
        m = idx.size
        poff = np.zeros(m + 1, dtype=np.int32)
        edges = np.empty(max(1, int(sub_off[-1])), dtype=_native.EDGE_DTYPE)
        scores = np.zeros(max(1, m), dtype=np.float64)
        status = np.zeros(max(1, m), dtype=np.int32)
        _native.check(lib.lt_tag_batch_host(batch, _native.ptr(sub_text), _native.ptr(sub_off32), m, W.beam, _native.ptr(poff),
                                            _native.ptr(edges), edges.size, _native.ptr(scores), _native.ptr(status)))
        return np.diff(poff), edges[:int(poff[m])], scores[:m], status[:m]

    def timed(fn, reps, together=True):
        fn()
        fn()
        torch.cuda.synchronize(dev)
        if together:
            dist.barrier()
        best = None
        for _ in range(reps):
            flush.fill_(1)
            torch.cuda.synchronize(dev)
            if together:
                dist.barrier()
            t0 = time.perf_counter()
            out = fn()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        return best, out

    reps = 3
    t_shard, local = timed(lambda: tag_indices(mine), reps)

    # gather of the packed results on rank 0 (sharding.gather_packed_contiguous: contiguous shards need no reordering
    # and no padding — every rank's arrays go point to point into their place in one device buffer per array, then
    # one copy per array into pinned host memory)
    pinned_cache = {}

    def gather():
        return sharding.gather_packed_contiguous(local, rank, world, device=dev, pinned=pinned_cache)

    gather()                              # (first use of the collective: communicator set-up is not part of a gather)
    torch.cuda.synchronize(dev)
    dist.barrier()
    t0 = time.perf_counter()
    gathered = gather()
    torch.cuda.synchronize(dev)
    dist.barrier()
    t_gather = time.perf_counter() - t0

    times = torch.tensor([t_shard], dtype=torch.float64, device=dev)
    all_times = [torch.zeros_like(times) for _ in range(world)]
    dist.all_gather(all_times, times)
    per_rank_ms = [1e3 * float(t[0]) for t in all_times]
    work = torch.tensor([float(lengths[mine].sum())], dtype=torch.float64, device=dev)
    all_work = [torch.zeros_like(work) for _ in range(world)]
    dist.all_gather(all_work, work)

    result = None
    if rank == 0:
        # the same corpus on ONE GPU (rank 0), same entry point: the N = 1 reference of the efficiency
        t_single, whole = timed(lambda: tag_indices(np.arange(n, dtype=np.int64)), reps, together=False)
        plen1, edges1, scores1, status1 = whole
        path_off, out_edges, scores_all, status_all = gathered
        identical = (np.array_equal(np.diff(path_off), plen1) and np.array_equal(out_edges, edges1) and
                     np.array_equal(scores_all, scores1) and np.array_equal(status_all, status1))
        t_total = max(per_rank_ms) * 1e-3 + t_gather
        imbalance = max(per_rank_ms) / (sum(per_rank_ms) / world)
        result = {
            'corpus': '%d %s-shaped sentences (chunks of 10k from seeds 7000+), built on rank 0 and broadcast; %d-morpheme dictionary, beam=%d'
                      % (n, config, W.cfg['n_dict'], W.beam),
            'sentences': int(n), 'tagged_ok': int((status_all == 0).sum()),
            't_single_gpu_ms': 1e3 * t_single, 't_shard_ms_max': max(per_rank_ms), 't_gather_ms': 1e3 * t_gather,
            't_total_ms': 1e3 * t_total, 'per_rank_ms': per_rank_ms,
            'value': n / t_total, 'value_without_gather': n / (max(per_rank_ms) * 1e-3), 'unit': UNIT,
            'speedup': t_single / (max(per_rank_ms) * 1e-3), 'speedup_with_gather': t_single / t_total,
            'efficiency': t_single / (max(per_rank_ms) * 1e-3) / world,
            'imbalance': imbalance, 'work_units_by_rank': [float(w[0]) for w in all_work],
            'partition': 'contiguous slices of equal total length (sharding.shard_bounds)',
            'limiter': ('gather on one host' if t_gather > max(per_rank_ms) * 1e-3 * (imbalance - 1.0) * 2 and t_gather > 0.1 * t_total
                        else 'imbalance between shards' if imbalance > 1.05 else 'per-call host overhead and launch tail of smaller shards'),
            'identical_to_single_gpu': bool(identical), 'corpus_build_seconds': build_s,
            'timing': 'wall clock around lt_tag_batch_host on host buffers (copies inside), best of %d, barrier before each' % reps,
        }
    if W is not W_headline:
        W.tagger.close()
        W.reg_tagger.close()
    return result


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path, all host cores."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from lattice_based_tagger_b200 import synth
    cfg, dictionary, sents = synth.build_workload(args.config, rank=0, n_sent=args.sentences)
    beam = cfg['beam'] if args.beam is None else args.beam
    kind = cpu_kind()
    tag_to_morphs, rules = dictionary.tag_to_morphs, dictionary.rules
    # sentences without any dictionary edge raise in the reference: replace them, as the GPU arm does
    tag_reg, oracle = _cpu_tagger('port', tag_to_morphs, rules, None, None)
    hit = [bool(oracle.lattice(s)) for s in sents]
    good = [i for i, h in enumerate(hit) if h]
    for i in range(len(sents)):
        if not hit[i]:
            sents[i] = sents[good[i % len(good)]]
    feature_dic, coef = reference_feature_inputs(args.config, oracle, feature_sample(sents), beam)
    items = list(feature_dic.keys())
    arm = CpuArm(args.config, items, coef, beam, kind=kind)
    # bounded sample per step so that the whole run takes a couple of minutes
    dt, _ = arm.run(sents[:min(len(sents), 32 * arm.cores)])
    rate = min(len(sents), 32 * arm.cores) / dt
    budget = max(1.0, min(12.0, 120.0 / (args.steps + args.warmup)))
    n_sample = int(max(arm.cores * 8, min(len(sents), rate * budget)))
    sample = sents[:n_sample]
    for _ in range(args.warmup):
        arm.run(sample)
    total = 0.0
    for _ in range(args.steps):
        dt, _ = arm.run(sample)
        total += dt
    arm.close()
    value = len(sample) * args.steps / total
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args.config, cfg, len(sents)),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': arm.cores, 'kind': arm.kind,
                         'sample': '%d of %d sentences per step, multiprocessing.Pool(%d) over %s'
                                   % (len(sample), len(sents), arm.cores, arm.describe())},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    emit(line)


def reference_feature_inputs(config, oracle, sample, beam):
    """Feature dictionary for the reference arm, drawn exactly as the GPU arm draws it but with the
    oracle as lattice / best-path provider (the two providers return identical results)."""
    import lattice_based_tagger_b200 as pkg
    from lattice_based_tagger_b200 import synth

    class _Seq:
        def __init__(self, hyp):
            self.sequences = [pkg.dictionary.Word(*w) for w in hyp.words]

    def tag_fn(sents):
        out = []
        for s in sents:
            try:
                out.append(_Seq(oracle.tag(s, beam)))
            except IndexError:
                out.append(None)
        return out

    def lattice_fn(sents):
        out = []
        for s in sents:
            edges = sorted(oracle.lattice(s), key=lambda w: (w[7], w[6]))     # device order
            words = [pkg.dictionary.Word(*w) for w in edges]
            bindex = []
            if words:
                bindex = [[] for _ in range(len(s.replace(' ', '')))]
                for w in words:
                    bindex[w.b].append(w)
            bos = pkg.dictionary.Word('BOS', 'BOS', None, 'BOS', None, 0, 0, 0, False)
            out.append(([bos] + words + [bos], bindex))
        return out

    cfg = synth.CONFIGS[config]
    tags = list(oracle.view.tag_to_morphs)
    vocab = None
    if cfg['n_feat'] > 1_000_000:
        vocab = [(m, t) for t, ms in oracle.view.tag_to_morphs.items() for m in sorted(ms)]
    return synth.make_features(sample, tag_fn, lattice_fn, cfg['n_feat'], tags, seed=3, vocab=vocab)


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
    # One string-hash seed for this process and every process it spawns: the order of the reference's
    # two-element conjugation set (lemmatizer.py:107) depends on it, and the GPU tables, the oracle and the
    # reference workers of one run must observe the same order.
    if os.environ.get('PYTHONHASHSEED') != '0':
        os.environ['PYTHONHASHSEED'] = '0'
        os.execv(sys.executable, [sys.executable] + sys.argv)
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', default='c2')
    ap.add_argument('--sentences', type=int, default=None, help='override the number of sentences per GPU')
    ap.add_argument('--beam', type=int, default=None)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-api', action='store_true')
    ap.add_argument('--other-configs', default='c3,c4,c5',
                    help='bounded samples of these configurations after the headline (N=1 only); "none" disables')
    ap.add_argument('--no-strong', action='store_true', help='N>1: skip the sharded-corpus (strong scaling) section')
    ap.add_argument('--strong-sentences', type=int, default=200_000)
    ap.add_argument('--strong-config', default='c3', help='configuration of the sharded corpus (BASELINE configs[2])')
    args = ap.parse_args()
    args.other_configs = [c for c in args.other_configs.split(',') if c and c != 'none' and c != args.config]
    capture_stdout()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == '__main__':
    main()
