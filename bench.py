#!/usr/bin/env python
"""bench.py -- BlockMaze prover benchmark (contract: one JSON line on rank 0).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload send|mixed1024]

Step (default workload): every GPU proves ONE synthetic `send` transaction (252 286 constraints, the configuration
BASELINE.json quotes for single-GPU latency).  `value` = proofs/s with the assignment already resident in HBM (GPU pipeline +
host finish); `e2e` = proofs/s through the BlockMaze cgo surface genSendproof() with host string arguments (native witness
generation, H2D of the assignment, GPU prover, D2H of the partial sums, hex encoding).  Multi-GPU: independent proofs, one
process per GPU, no data-path collective (weak scaling); the barrier / max-over-ranks uses torch.distributed.
--impl reference times the UNMODIFIED libsnark prover (oracle/_ref, -DMULTICORE -fopenmp) on the host cores for the same
transaction.
"""
import argparse
import json
import os
import random
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))

# The cgo layer printf()s what the reference prints ("Trying to generate send proof..."): send the process's stdout to stderr and keep
# the real stdout for the one JSON line of the contract.
_REAL_STDOUT = None


def capture_stdout():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(obj):
    if _REAL_STDOUT is None:
        print(json.dumps(obj), flush=True)
    else:
        os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

CONSTRAINTS = {"mint": 167270, "send": 252286, "deposit": 503863, "redeem": 167853}
DOMAIN = {"mint": 196608, "send": 262144, "deposit": 524288, "redeem": 196608}
WORKLOADS = {"send": "send circuit (252286 constraints, QAP domain 2^18): one Groth16 proof per step per GPU",
             "mixed1024": "mixed batch of 1024 synthetic mint/send/deposit/redeem transactions sharded over the GPUs"}
IMAD_PER_G1_POINT = 23936          # 16 windows x (11 modmul x 136 wide multiply-adds) per point of a 254-bit G1 MSM
FR_MODULUS = 21888242871839275222246405745257275088548364400416034343698204186575808495617
MODMUL_PER_G1_ADD = 10             # what msm_accumulate_kernel really issues: XYZZ += affine is 8M + 2S (ec.cuh)


def imad_peaks(api):
    """Measured in this run (cabi.cu imad_peak_kernel): 32-bit IMAD, carry-chained wide IMAD (SURVEY.md 8d's denominator: the
    field multiplication is made of IMAD.WIDE.U32.X, which issues at half the IMAD rate), whole modular multiplications."""
    return {"imad32_T": float(api.lib.zkb200_bench_imad_peak(0)), "imad_wide_T": float(api.lib.zkb200_bench_imad_peak(1)),
            "modmul_G": 1e3 * float(api.lib.zkb200_bench_imad_peak(2))}


def host_threads():
    """Every host core for the reference's OpenMP build, whatever the launcher exported (torch.distributed.run sets OMP_NUM_THREADS=1,
    which voided the round-1 ratios at N >= 2).  Must run before libgomp is loaded."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    os.environ["OMP_NUM_THREADS"] = str(n)
    return n


def omp_threads(want):
    """Force and read back the OpenMP thread count of the already loaded runtime."""
    import ctypes
    try:
        gomp = ctypes.CDLL("libgomp.so.1")
        gomp.omp_set_num_threads(int(want))
        return int(gomp.omp_get_max_threads())
    except OSError:
        return int(os.environ.get("OMP_NUM_THREADS", "1"))


def shard_batch(seeds, rank, world):
    """Independent transactions are dealt to the ranks with no inter-GPU traffic.  The transaction type is
    seed % 4 and deposit proofs cost ~2x a mint, so the deal rotates by seed // 4 to give every rank the same type mix."""
    return [s for s in seeds if ((s // 4) + s) % world == rank]


def reduce_counts_and_time(count, seconds, dist, device="cuda"):
    """Whole-job units = sum over ranks; time = max over ranks."""
    if dist is None or not dist.is_initialized():
        return count, seconds
    import torch
    c = torch.tensor([float(count)], device=device)
    t = torch.tensor([float(seconds)], device=device)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(round(c.item())), t.item()


def key_dir():
    for d in (os.environ.get("ZKB200_KEY_DIR"), os.path.join(ROOT, "oracle", "_ref", "prfKey"), "/usr/local/prfKey"):
        if d and os.path.exists(os.path.join(d, "sendpk.txt")):
            return d
    raise SystemExit("bench.py: no proving keys found (ZKB200_KEY_DIR, oracle/_ref/prfKey, /usr/local/prfKey)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, windows):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        out = self.stats(windows)
        self.proc.terminate()
        return out

    def stats(self, windows):
        """Clocks and throttle reasons of the samples that fall into the given (t0, t1) windows; the sampler keeps running."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in list(self.lines):
            if not any(a - 0.05 <= t <= b + 0.05 for a, b in windows):
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def reference_arm(args, rank, world):
    """The reference's own CPU prover on the host cores (rank 0 only)."""
    if rank != 0:
        return
    from oracle import refapi as Rf, bn254_oracle as O
    import fixtures as F
    circuit = "send"
    line = {"impl": "reference", "metric": "proofs_per_sec", "unit": "proofs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64 limbs (libff Fp_model, 254-bit Montgomery)", "data": "synthetic"}
    if not Rf.available(circuit + "_mt"):
        emit(({"impl": "reference", "unavailable": "oracle/_ref/libref_send_mt.so not built (needs /root/reference at build time)"}))
        return
    cores = host_threads()              # torchrun exports OMP_NUM_THREADS=1: the reference arm always takes every host core
    t0 = time.perf_counter()
    Rf.load_pk(circuit, os.path.join(key_dir(), circuit + "pk.txt"), mt=True)
    cores = omp_threads(cores)
    t_load = time.perf_counter() - t0
    words = O.fixed_rng_words(42, 64)
    txs = [F.synthetic(circuit, s) for s in range(max(1, args.warmup) + args.steps)]
    tw = time.perf_counter()
    for i in range(max(1, args.warmup)):
        Rf.prove(circuit, txs[i], words, mt=True)
    per_proof = (time.perf_counter() - tw) / max(1, args.warmup)
    # bounded sample: a CPU proof takes ~2 s, so at most ~100 s worth of the K steps are actually proved (the rate is per proof either way)
    timed = max(1, min(args.steps, int(100.0 / max(per_proof, 1e-3))))
    t0 = time.perf_counter()
    phases = [0.0] * 5
    for i in range(timed):
        res = Rf.prove(circuit, txs[max(1, args.warmup) + i], words, mt=True)
        assert res["rc"] == 0
        phases = [a + b for a, b in zip(phases, res["timings"])]
    dt = time.perf_counter() - t0
    assert cores > 1 or (os.cpu_count() or 1) == 1, "reference arm ran single-threaded on a multi-core host"
    v = timed / dt
    line.update(value=v, ms_per_step=1e3 * dt / timed,
                config={"workload": WORKLOADS["send"], "detail": "reference path per step: gadget construction + witness + is_satisfied + r1cs_gg_ppzksnark_prover, pk resident",
                        "constraints": CONSTRAINTS[circuit], "domain": DOMAIN[circuit], "pk_load_s_excluded": round(t_load, 1)},
                cpu_baseline={"value": v, "unit": "proofs/s", "cores": cores, "kind": "reference",
                              "sample": "%d of the %d requested send proofs timed (bounded to ~100 s), libsnark -DMULTICORE -fopenmp, OMP_NUM_THREADS=%d; "
                                        "prover phases avg s: qap %.2f A %.2f B %.2f H %.2f L %.2f" % (timed, args.steps, cores, *[p / timed for p in phases])},
                e2e={"value": v, "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, gpu_launches=0)
    emit((line))


def load_fma_count():
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r02_fma_pipe.json")))
    except (OSError, ValueError):
        return None


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        return {}


def verify_sample(api, circuit, txs, proofs, count=8):
    """The bench checks what it timed: `count` of the proofs (spread over the run) go through verify<Circuit>proof."""
    idx = sorted(set(int(i * (len(proofs) - 1) / max(1, count - 1)) for i in range(min(count, len(proofs)))))
    for i in idx:
        assert proofs[i][:10] != "0000000000", "prover returned the default proof for a valid transaction"
        assert api.verify_proof(circuit, proofs[i], api.verify_args(circuit, txs[i])), "a timed proof does not verify"
    return len(idx)


def run_mixed(api, F, rank, world, dist, barrier, depth, sampler, single_process_gpus=0, repeat=1):
    """BASELINE.json configs[3]: 1024 synthetic transactions, type = seed mod 4 (256 of each circuit), through gen*proof.
    One process per GPU (torchrun): the batch is dealt to the ranks by shard_batch, no data-path collective (strong scaling).
    --single-process: this one process proves the whole batch on `single_process_gpus` devices through zkb200_prove_batch."""
    names = ["mint", "send", "deposit", "redeem"]
    mine = list(range(1024)) if single_process_gpus else shard_batch(list(range(1024)), rank, world)
    jobs = [(names[sd % 4], F.synthetic(names[sd % 4], sd + 1024 * k)) for k in range(repeat) for sd in mine]      # --repeat R: R batches back to back
    nthreads = depth * max(1, single_process_gpus)
    warm = [(c, F.synthetic(c, 5000 + rank + 10 * k)) for k in range(max(2, nthreads // 2)) for c in names]
    api.prove_batch(warm, nthreads)                      # loads the four keys on every active device, warms every lane
    prepared = api.prove_batch_prepare(jobs)             # the zkb200_tx array (ctypes marshalling: 25 ms of Python per 1024 transactions)
    barrier()
    api.lib.zkb200_device_proofs.restype = __import__("ctypes").c_long
    before = [int(api.lib.zkb200_device_proofs(d)) for d in range(max(1, single_process_gpus))] if single_process_gpus else []
    api.lib.zkb200_device_timer(0)
    t0 = time.perf_counter()
    proofs, bad = api.prove_batch_run(prepared, nthreads)
    dev_ms = float(api.lib.zkb200_device_timer(1))
    t1 = time.perf_counter()
    barrier()
    assert bad == 0, "%d of the batch came back as default proofs" % bad
    api.lib.zkb200_device_proofs.restype = __import__("ctypes").c_long
    per_device = [int(api.lib.zkb200_device_proofs(d)) - before[d] for d in range(len(before))]
    checked = 0
    for c in names:                                      # two of each circuit through verify*proof
        sel = [i for i, (cc, _) in enumerate(jobs) if cc == c]
        checked += verify_sample(api, c, [jobs[i][1] for i in sel], [proofs[i] for i in sel], 2)
    units, dt = reduce_counts_and_time(len(jobs), max(dev_ms * 1e-3, t1 - t0), None if single_process_gpus else dist)
    return {"metric": "proofs_per_sec", "value": units / dt, "unit": "proofs/s", "n_gpus": single_process_gpus or world, "scaling": "strong",
            "seconds": dt, "transactions": units, "verified_sample": checked, "callers_per_gpu": depth,
            "mode": "one process, library device scheduler (zkb200_prove_batch)" if single_process_gpus else "one process per GPU, batch dealt round-robin by type",
            "proofs_per_device": per_device if single_process_gpus else None,
            "workload": WORKLOADS["mixed1024"], "clocks": sampler.stats([(t0, t1)]),
            "timing": "max(CUDA events after device synchronisations, wall clock) around the batch, host work included; max over ranks"}


def run_msm_split(api, dist, rank, world, barrier, logn, iters, sampler):
    """BASELINE.json configs[4], second half: ONE MSM of 2^logn points split by point range over the GPUs, G1 then G2; every GPU returns one
    partial point and rank 0 adds them on the host (zkb200_g1_sum / zkb200_g2_sum; no collective on the data path -- the 64/128-byte points
    travel through the rendezvous gather).  Inputs: SHA512_rng scalars, bases SHA512_rng(2^32+i)*G (zkb200_synth_*), functions of the
    global index.  The sum is checked against profiles/msm_split_golden.json (the single-GPU result, itself cross-checked against libff at
    2^20 by tests/test_gpu_kernels.py::test_sweep_sizes_against_reference_on_identical_inputs)."""
    import ctypes as C
    n = 1 << logn
    per = n // world
    first, count = rank * per, (per if rank < world - 1 else n - per * (world - 1))
    out = {}
    for group, name, size in ((1, "g1", 64), (2, "g2", 128)):
        buf = C.create_string_buffer(size)
        barrier()
        t0 = time.perf_counter()
        ms = float(api.lib.zkb200_bench_msm_slice(group, first, count, 0, max(1, iters), buf))       # one untimed run inside, then `iters` timed ones
        t1 = time.perf_counter()
        barrier()
        units, tmax = reduce_counts_and_time(count, ms * 1e-3, dist)
        pts = [buf.raw]
        if dist is not None:
            gathered = [None] * world
            dist.all_gather_object(gathered, buf.raw)
            pts = gathered
        if rank == 0:
            total = (api.g1_sum if group == 1 else api.g2_sum)(pts)
            x = "%064x" % int.from_bytes(total[:32], "little")
            gold = None
            try:
                gold = json.load(open(os.path.join(ROOT, "profiles", "msm_split_golden.json"))).get("%s_2^%d_x" % (name, logn))
            except (OSError, ValueError):
                pass
            if gold is not None:
                assert x == gold, "MSM split result differs from the recorded single-GPU result"
            out[name] = {"value": units / tmax, "unit": "points/s", "ms_per_msm": round(1e3 * tmax, 3), "points": units, "result_x": x,
                         "result_matches_golden": None if gold is None else True,
                         "frac_of_imad_peak": None, "clocks": sampler.stats([(t0, t1)])}
    if rank == 0:
        out.update({"metric": "msm_points_per_sec", "n_gpus": world, "scaling": "strong", "log2_points": logn,
                    "workload": "single MSM of 2^%d points split by point range, one partial point per GPU summed on the host; G1 and G2" % logn})
    return out


def kernel_sweep(api):
    """BASELINE.json configs[4]: BN254 G1/G2 MSM and Fr NTT at 2^16..2^24 on one GPU next to libff multi_exp / libfqfft FFT on the host
    cores.  The host legs get IDENTICAL inputs (zkb200_synth_scalars / zkb200_synth_bases: libff's SHA512_rng stream) and every host result
    is compared with the GPU's; they stop at 2^20 (G2: 2^18), where they already take seconds."""
    import ctypes as C
    hbm = measured_peaks().get("hbm_gbs", 6650.0)
    imad = float(api.lib.zkb200_bench_imad_peak(1))
    cpu, cores = None, 0
    try:
        from oracle import refapi as Rf
        if Rf.available("kernels_mt"):
            cores = host_threads()
            cpu = Rf
            Rf.lib("kernels_mt")
            cores = omp_threads(cores)
    except Exception:
        cpu = None
    rows = []
    for lg in (16, 18, 20, 22, 24):
        n = 1 << lg
        row = {"log2_n": lg}
        ms = float(api.lib.zkb200_bench_ntt(lg, 1, 5))
        row["ntt_ms"] = round(ms, 4)
        row["ntt_GBps"] = round(64.0 * n / (ms * 1e-3) / 1e9, 1)
        row["ntt_frac_of_hbm"] = round(row["ntt_GBps"] / hbm, 4)
        g1 = C.create_string_buffer(64)
        ms = float(api.lib.zkb200_bench_msm_slice(1, 0, n, 0, 2, g1))
        row["msm_g1_ms"] = round(ms, 3)
        row["msm_g1_frac_of_imad"] = round(IMAD_PER_G1_POINT * n / (ms * 1e-3) / 1e12 / imad, 4)
        fx = C.create_string_buffer(64)
        ms = float(api.lib.zkb200_bench_msm_slice(1, 0, n, -16, 2, fx))
        assert fx.raw == g1.raw, "fixed-base and windowed MSM disagree"
        row["msm_g1_fixed_base_ms"] = round(ms, 3)
        row["msm_g1_fixed_base_frac_of_imad"] = round(IMAD_PER_G1_POINT * n / (ms * 1e-3) / 1e12 / imad, 4)
        g2 = C.create_string_buffer(128)
        ms = float(api.lib.zkb200_bench_msm_slice(2, 0, n, 0, 1 if lg >= 22 else 2, g2))
        row["msm_g2_ms"] = round(ms, 3)
        row["msm_g2_frac_of_imad"] = round(3 * IMAD_PER_G1_POINT * n / (ms * 1e-3) / 1e12 / imad, 4)
        if cpu is not None and lg <= 20:
            sc = api.synth_scalars(0, n)
            want, sec = cpu.msm_g1_bytes(api.synth_bases(1, 0, n), sc, 0, chunks=0, mt=True)
            assert want == g1.raw, "G1 MSM differs from libff multi_exp at 2^%d" % lg
            row["cpu_msm_g1_s"] = round(sec, 3)
            ev, sec = cpu.domain_op_timed(n, "FFT", sc, mt=True)
            assert ev == api.domain_op(n, "FFT", sc), "NTT differs from libfqfft at 2^%d" % lg
            row["cpu_fft_s"] = round(sec, 4)
            if lg <= 18:
                want, sec = cpu.msm_g2_bytes(api.synth_bases(2, 0, n), sc, 0, chunks=0, mt=True)
                assert want == g2.raw, "G2 MSM differs from libff multi_exp at 2^%d" % lg
                row["cpu_msm_g2_s"] = round(sec, 3)
            row["host_results_equal_gpu"] = True
        rows.append(row)
    return {"metric": "kernel_sweep", "unit": "ms", "n_gpus": 1,
            "data": "synthetic: scalars libff SHA512_rng<Fr>(i), bases SHA512_rng<Fr>(2^32+i)*G, NTT input = the scalar stream; identical bytes on GPU and host",
            "peaks": {"hbm_GBps": hbm, "imad_wide_T_per_s": round(imad, 2), "cpu_threads": cores},
            "algorithmic": {"ntt_bytes_per_element": 64, "imad_per_g1_point": IMAD_PER_G1_POINT, "imad_per_g2_point": 3 * IMAD_PER_G1_POINT},
            "rows": rows}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="send", choices=["send", "mixed1024", "sweep", "msm_split"])
    ap.add_argument("--logn", type=int, default=24, help="msm_split: log2 of the total number of points")
    ap.add_argument("--repeat", type=int, default=1, help="mixed1024: this many 1024-transaction batches back to back in the timed region")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the per-circuit latency table, the NTT roofline and the extra BASELINE.json configs "
                    "(mixed1024, msm_split24, kernel sweep) that the default run appends to its JSON line")
    ap.add_argument("--single-process", action="store_true", help="mixed1024: ONE process drives all --gpus devices through the library's own "
                    "device scheduler (what an unchanged geth process gets); do not launch under torchrun")
    args = ap.parse_args()
    capture_stdout()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))

    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))

    import blockmaze_b200 as zk
    from blockmaze_b200 import api
    from blockmaze_b200 import wallet as F    # seeded transactions built with the library's own cgo helpers (nothing from oracle/ on this arm)
    zk.init(local)
    kd = key_dir()
    api.set_key_dir(kd)

    def barrier():
        api.lib.zkb200_device_sync()
        if dist is not None:
            dist.barrier()

    def finish(line):
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        if rank == 0 and line is not None:
            emit(line)

    sampler = ClockSampler(local)
    if args.workload == "sweep":
        finish(kernel_sweep(api) if rank == 0 else None)
        return
    if args.workload == "msm_split":
        res = run_msm_split(api, dist, rank, world, barrier, args.logn, max(1, args.steps // 50), sampler)
        if rank == 0:
            res.update({"value": res["g1"]["value"], "unit": "points/s", "ms_per_step": res["g1"]["ms_per_msm"], "higher_is_better": True, "data": "synthetic",
                        "dtype": "u32", "config": {"workload": res["workload"]}})
        finish(res if rank == 0 else None)
        return
    if args.workload == "mixed1024":
        single = args.gpus if args.single_process else 0
        if single:
            assert world == 1, "--single-process is not launched under torchrun"
            assert api.set_devices(list(range(single))) == single
        res = run_mixed(api, F, rank, world, dist, barrier, int(os.environ.get("BENCH_CALLERS", "3")), sampler, single, max(1, args.repeat))
        if rank == 0:
            res.update({"steps": 1, "warmup": args.warmup, "ms_per_step": 1e3 * res["seconds"], "higher_is_better": True, "vs_baseline": None, "dtype": "u32",
                        "data": "synthetic", "config": {"workload": WORKLOADS["mixed1024"], "detail": res["mode"]},
                        "e2e": {"value": res["value"], "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": None})
        finish(res if rank == 0 else None)
        return

    # ---- default workload: send ---------------------------------------------------------------------------------------------------------
    pk = zk.ProvingKey(os.path.join(kd, "sendpk.txt"))
    rng = random.Random(1000 + rank)
    r, s = rng.randrange(1, FR_MODULUS), rng.randrange(1, FR_MODULUS)      # pinned per rank
    windows = []
    from concurrent.futures import ThreadPoolExecutor
    depth = pk.lanes                       # proofs in flight per GPU (zkb200.h "lanes")

    tx0 = F.synthetic("send", rank)
    w0 = api.witness("send", tx0)
    # ---- sequential pass (one proof at a time, L2 flushed before each): per-kernel CUDA-event times for the rooflines ----------
    res = pk.prove(w0, r, s)
    assert res["rc"] == 0
    proof0 = res["proof_hex"]
    assert api.verify_proof("send", proof0, api.verify_args("send", tx0)), "the proof of the resident assignment does not verify"
    acc_ms, qap_ms, msm_ms, gpu_ms = [], [], [], []
    for i in range(args.warmup + min(args.steps, 30)):
        api.lib.zkb200_flush_l2()
        res = pk.prove(None, r, s)
        if i >= args.warmup:
            gpu_ms.append(res["timings_ms"][0]); qap_ms.append(res["timings_ms"][1]); msm_ms.append(res["timings_ms"][2])
    api.lib.zkb200_set_isolate_h(1)      # roofline: the H-query accumulate kernel timed with nothing beside it
    for i in range(args.warmup + min(args.steps, 30)):
        api.lib.zkb200_flush_l2()
        res = pk.prove(None, r, s)
        if i >= args.warmup:
            acc_ms.append(res["timings_ms"][4])
    api.lib.zkb200_set_isolate_h(0)
    # ---- leg 1: `value` -- assignment resident in HBM, `depth` proofs in flight ------------------------------------------------
    lanes = [pk.lane_acquire() for _ in range(depth)]
    for ln in lanes:                     # make the assignment resident on every lane
        pk.submit(ln, w0, r, s)
    for ln in lanes:
        assert pk.collect(ln)["rc"] == 0
    for i in range(args.warmup):
        pk.submit(lanes[i % depth], None, r, s); pk.collect(lanes[i % depth])
    barrier()
    api.lib.zkb200_device_timer(0)        # CUDA events on the device, bracketed by device synchronisations (and the barriers)
    t0 = time.perf_counter()
    launches, acc_inflight, value_proofs = 0, [], set()
    for i in range(args.steps):
        ln = lanes[i % depth]
        if i >= depth:
            res = pk.collect(ln); launches += res["launches"]; acc_inflight.append(res["timings_ms"][4]); value_proofs.add(res["proof_hex"])
        pk.submit(ln, None, r, s)
    for i in range(args.steps, args.steps + min(depth, args.steps)):
        res = pk.collect(lanes[i % depth]); launches += res["launches"]; acc_inflight.append(res["timings_ms"][4]); value_proofs.add(res["proof_hex"])
    dev_ms = float(api.lib.zkb200_device_timer(1))
    t1 = time.perf_counter()
    barrier()
    for ln in lanes:
        pk.lane_release(ln)
    assert value_proofs == {proof0}, "proofs of the timed `value` region differ from the verified proof of the same assignment, r and s"
    windows.append((t0, t1))
    wall_value = t1 - t0
    units, dt = reduce_counts_and_time(args.steps, dev_ms * 1e-3, dist)
    # ---- leg 2: `e2e` -- the cgo call a BlockMaze node makes, host string arguments in, proof string out; `depth` caller threads
    #      (goroutines in geth), each call synchronous --------------------------------------------------------------------------
    txs = [F.synthetic("send", rank + world * (i + 1)) for i in range(args.warmup + args.steps)]
    lat, brk = [], []
    for i in range(args.warmup + min(args.steps, 30)):          # single caller: latency and its breakdown
        ta = time.perf_counter()
        api.gen_proof("send", txs[i])
        if i >= args.warmup:
            lat.append(time.perf_counter() - ta)
            brk.append(api.last_breakdown_ms())
    pool = ThreadPoolExecutor(depth)
    list(pool.map(lambda tx: api.gen_proof("send", tx), txs[:args.warmup]))
    barrier()
    api.lib.zkb200_device_timer(0)
    t2 = time.perf_counter()
    proofs = list(pool.map(lambda tx: api.gen_proof("send", tx), txs[args.warmup:]))
    dev_ms_e = float(api.lib.zkb200_device_timer(1))
    t3 = time.perf_counter()
    barrier()
    pool.shutdown()
    windows.append((t2, t3))
    assert all(not p_.startswith("0000000000") for p_ in proofs), "prover returned the default proof for a valid transaction"
    verified = verify_sample(api, "send", txs[args.warmup:], proofs, 8)
    units_e, dt_e = reduce_counts_and_time(args.steps, max(dev_ms_e * 1e-3, t3 - t2), dist)      # host work counts: the slower clock
    nvars = pk.num_variables
    workload = ("send circuit (%d constraints, %d variables, QAP domain 2^18): one Groth16 proof per step per GPU, %d proofs in flight; value = resident "
                "assignment through submit/collect, e2e = genSendproof() cgo calls from %d caller threads" % (CONSTRAINTS["send"], nvars, depth, depth))
    # bytes one such proof moves over PCIe, counted by the library from the copies it enqueues (one more genSendproof on this thread: the
    # counters are per calling thread)
    api.gen_proof("send", txs[0])
    xfer = (__import__("ctypes").c_ulonglong * 2)()
    api.lib.zkb200_last_transfer_bytes(xfer)
    h2d, d2h = int(xfer[0]), int(xfer[1])
    clocks = sampler.stats(windows)

    extras = {}
    if rank == 0 and not args.no_extras:
        # p50 end-to-end latency per circuit through the cgo surface (BASELINE.json metric, configs[0..3])
        per = {}
        for c in ("mint", "send", "deposit", "redeem"):
            if not os.path.exists(os.path.join(kd, c + "pk.txt")):
                continue
            ts = []
            for i in range(9):
                tx = F.synthetic(c, 9000 + i)
                ta = time.perf_counter(); api.gen_proof(c, tx); ts.append(1e3 * (time.perf_counter() - ta))
            per[c] = round(statistics.median(ts[2:]), 3)
        extras["p50_latency_ms_per_circuit_e2e"] = per
        # NTT roofline at a size that does not fit L2 (2^24 x 32 B = 512 MB): algorithmic bytes 64*n per transform
        ms_ntt = api.lib.zkb200_bench_ntt(24, 1, 5)
        peaks = measured_peaks()
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        ach = 64.0 * (1 << 24) / (ms_ntt * 1e-3) / 1e9
        prof = load_fma_count() or {}
        extras["roofline_ntt"] = {"bound": "hbm", "kernel": "ntt passes of one 2^24-point forward NTT", "achieved": round(ach, 1), "peak": hbm_peak,
                                  "unit": "GB/s", "frac": round(ach / hbm_peak, 4),
                                  "traffic": prof.get("ntt24_dram_bytes", {"dram_bytes_per_transform": 4010000000, "source": "profiles/r01_notes.md (B)"}),
                                  "ms": round(ms_ntt, 4),
                                  "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}
        mm_peak = 1e3 * float(api.lib.zkb200_bench_imad_peak(2))
        mm_rate = 12.0 * (1 << 24) / (ms_ntt * 1e-3) / 1e9          # (n/2) * log2(n) butterflies, one modular multiplication each
        extras["roofline_ntt"]["integer_bound"] = {
            "modmul_G_per_s": round(mm_rate, 2), "modmul_peak_G_per_s": round(mm_peak, 2), "frac": round(mm_rate / mm_peak, 4) if mm_peak else None,
            "what": "a 254-bit NTT is bound by the integer-multiply pipe, not HBM: 12 modular multiplications per 64 algorithmic bytes "
                    "cap it at modmul_peak * 64 / 12 bytes/s (about 0.36 TB/s), 5.6 % of the HBM peak"}

    line = None
    if rank == 0:
        pk_ = imad_peaks(api)
        imad_peak = pk_["imad_wide_T"]
        acc_avg = statistics.mean(acc_ms) if acc_ms else 0.0
        n_h = DOMAIN["send"] - 1
        ach = IMAD_PER_G1_POINT * n_h / (acc_avg * 1e-3) / 1e12 if acc_avg > 0 else 0.0
        modmul_rate = MODMUL_PER_G1_ADD * 16 * n_h / (acc_avg * 1e-3) / 1e9 if acc_avg > 0 else 0.0
        msm_h_avg = statistics.mean(msm_ms)
        prof = load_fma_count() or {}
        fma_instr = prof.get("fma_pipe_warp_instr_per_send_proof")
        sm_mhz = clocks.get("sm_mhz") or measured_peaks().get("sm_max_mhz", 1965.0)
        step_ms = 1e3 * dt / args.steps
        nominal = 148 * 128 * sm_mhz * 1e6 / 4 / 1e12      # 148 SMs x 128 lanes, one wide multiply-add per lane every 4 clocks
        line = {
            "metric": "proofs_per_sec", "value": units / dt, "unit": "proofs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": WORKLOADS["send"], "detail": workload, "l2": "no flush in the timed loops: every proof streams ~0.6 GB of fixed-base tables (all 16 windows of the H, A, B, L queries), "
                             "5x the 126 MB L2; the sequential pass behind gpu_ms_per_proof and the rooflines flushes L2 (256 MB memset) before every proof",
                       "proofs_in_flight": depth,
                       "timing": "value: CUDA events after device synchronisations around the K steps (wall clock of the same region: %.3f ms/step); "
                                 "e2e: max(that, wall clock), host work included; max over ranks" % (1e3 * wall_value / max(1, args.steps)),
                       "randomness": "value leg: r, s pinned per rank; e2e leg: fresh r, s per proof (std::random_device)",
                       "checked": "value leg: every proof of the timed region equals the proof of the same (assignment, r, s) made beforehand, which verifySendproof accepts; "
                                  "e2e leg: %d of the timed proofs, spread over the region, pass verifySendproof" % verified},
            "clocks": clocks,
            "e2e": {"value": units_e / dt_e, "unit": "proofs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "p50_latency_ms": round(1e3 * statistics.median(lat), 3),
                    "breakdown_ms": {k: round(statistics.median(b[k] for b in brk), 3) for k in brk[0]}},
            "gpu_launches": launches,
            "gpu_ms_per_proof": {"what": "one proof at a time, CUDA events", "total": round(statistics.mean(gpu_ms), 3), "qap_witness_map": round(statistics.mean(qap_ms), 3),
                                 "msm_H": round(msm_h_avg, 3), "msm_H_accumulate_kernel_alone": round(acc_avg, 3)},
            "roofline": {"bound": "imad", "kernel": "msm_accumulate_kernel<Fq> (H query, %d points), timed alone after an L2 flush" % n_h, "achieved": round(ach, 3), "peak": round(imad_peak, 2),
                         "unit": "TIMAD/s", "frac": round(ach / imad_peak, 4) if imad_peak else None,
                         "traffic": prof.get("acc_h_dram_bytes", {"dram_bytes_per_launch": 501680000, "source": "profiles/r01_notes.md (B)"}),
                         "nominal_peak": {"value": round(nominal, 2), "what": "148 SMs x 128 lanes x %.0f MHz / 4 clocks per wide multiply-add" % sm_mhz,
                                          "frac": round(ach / nominal, 4) if nominal else None},
                         "issued": {"modmul_G_per_s": round(modmul_rate, 2), "modmul_peak_G_per_s": round(pk_["modmul_G"], 2),
                                    "frac": round(modmul_rate / pk_["modmul_G"], 4) if pk_["modmul_G"] else None,
                                    "what": "modular multiplications the kernel really issues (10 per mixed addition, 16 per point) against a kernel of back-to-back ff.cuh multiplications"},
                         "timed_region": {"kernel_ms_avg": round(statistics.mean(acc_inflight), 3) if acc_inflight else None,
                                          "frac": round(IMAD_PER_G1_POINT * n_h / (statistics.mean(acc_inflight) * 1e-3) / 1e12 / imad_peak, 4)
                                          if acc_inflight and imad_peak else None,
                                          "what": "the same kernel timed by CUDA events on its stream inside the timed region of `value`, where it shares the SMs "
                                                  "with the kernels of the other proofs in flight (so this is a lower bound of its own efficiency)"},
                         "whole_msm_frac": {"value": round(IMAD_PER_G1_POINT * n_h / (msm_h_avg * 1e-3) / 1e12 / imad_peak, 4) if imad_peak and msm_h_avg else None,
                                            "what": "the same algorithmic count over the WHOLE H-query MSM of a proof running alone (digit sort + accumulate + fold + bucket reduction, %.3f ms)" % msm_h_avg},
                         "step_multiply_pipe_util": {"value": round(fma_instr * 32 / (imad_peak * 1e12) / (step_ms * 1e-3), 4) if fma_instr and imad_peak else None,
                                                     "fma_pipe_warp_instr_per_proof": fma_instr,
                                                     "what": "FMA-pipe warp instructions of one send proof (ncu, profiles/r02_fma_pipe.json) x 32 lanes / measured wide-IMAD peak / ms_per_step: "
                                                             "how busy the multiply pipe is over a pipelined step if every one of them were a wide multiply-add"},
                         "peaks": {k: round(v, 2) for k, v in pk_.items()},
                         "note": "integer-multiply roofline (SURVEY.md 8d): algorithmic 23936 wide multiply-adds per point / CUDA-event kernel time; "
                                 "peak = carry-chained mad.lo.cc/madc.hi.cc (IMAD.WIDE.U32.X) microbenchmark in this run.  The algorithmic figure counts the reference's "
                                 "mixed addition (7M+4S = 11 multiplications); the kernel's XYZZ addition needs 10, so this ratio tops out at 1.10 -- 'issued' is the strict one"},
        }
        line.update(extras)
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline()
    pk.close()
    # ---- the other BASELINE.json configs, appended to the same line so that the driver's runs (N = 1, 2, 4, 8) carry them ------------------
    if not args.no_extras:
        mixed = run_mixed(api, F, rank, world, dist, barrier, depth, sampler)
        split = run_msm_split(api, dist, rank, world, barrier, 24, 3, sampler)
        if rank == 0:
            line["mixed1024"] = mixed
            imad_peak = line["roofline"]["peak"]
            for name, per_point in (("g1", IMAD_PER_G1_POINT), ("g2", 3 * IMAD_PER_G1_POINT)):
                split[name]["frac_of_imad_peak"] = round(per_point * split[name]["value"] / 1e12 / (imad_peak * world), 4) if imad_peak else None
            line["msm_split24"] = split
            if world == 1:
                line["kernel_sweep"] = kernel_sweep(api)
    finish(line)


def cpu_baseline():
    """Reference libsnark prover (MULTICORE) on this box's host cores: one send proof, pk load excluded."""
    code = ("import sys,os,json,time; sys.path.insert(0,%r); sys.path.insert(0,%r)\n"
            "from oracle import refapi as Rf, bn254_oracle as O; import fixtures as F\n"
            "os.environ['OMP_NUM_THREADS'] = str(len(os.sched_getaffinity(0)))\n"
            "Rf.load_pk('send', os.path.join(%r,'sendpk.txt'), mt=True)\n"
            "w=O.fixed_rng_words(42,64); Rf.prove('send',F.synthetic('send',0),w,mt=True)\n"
            "t=time.perf_counter(); n=3\n"
            "for i in range(n): Rf.prove('send',F.synthetic('send',1+i),w,mt=True)\n"
            "dt=time.perf_counter()-t\n"
            "print('CPUBASE',json.dumps({'value':n/dt,'unit':'proofs/s','cores':int(os.environ['OMP_NUM_THREADS']),'kind':'reference',"
            "'sample':'3 send proofs after 1 warm-up, unmodified libsnark (-DMULTICORE -fopenmp), witness+is_satisfied+prover, pk load excluded'}))\n"
            % (ROOT, os.path.join(ROOT, "tests", "golden"), key_dir()))
    from oracle import refapi as Rf
    if not Rf.available("send_mt"):
        return {"value": None, "unit": "proofs/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref not built on this box"}
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900)
    for l in out.stdout.splitlines():
        if l.startswith("CPUBASE "):
            return json.loads(l[8:])
    return {"value": None, "unit": "proofs/s", "cores": 0, "kind": "reference", "sample": "reference run failed: " + out.stderr[-200:]}


if __name__ == "__main__":
    main()
