#!/usr/bin/env python
"""bench.py -- CPHF polarizability hot path (AO2MO + Hessian assembly + multi-frequency solve).

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K ...   # reference algorithm on host cores

Metric (BASELINE.json): FP64 TFLOP/s of the end-to-end CPHF property calculation, counted with
the algorithmic flop model of SURVEY.md section 8(d) (never the flops actually executed), plus
wall time per step (``ms_per_step``).  One step = one full property calculation:
AO->MO transform of the resident AO ERI tensor, A/B assembly, every frequency solved, property
tensors formed.

Workloads (``config.workload``; ``--workload`` selects another one as the main record):
  config5 (default, EVERY N): synthetic RHF nbasis=256 nocc=32: magnetizability (static, imaginary
           operator) + 64-frequency polarizability sweep.  The largest config whose dense AO tensor
           (34.4 GB) fits one B200; for N > 1 the SAME job is strong-scaled (AO tensor sharded over
           nu, frequencies dealt out), so the 1 -> 8 curve is one workload.
  config3 : synthetic RHF nbasis=512 nocc=64, polarizability at 8 frequencies, AO tensor sharded
           8 ways (68.7 GB per GPU) -- BASELINE's headline; needs 8 GPUs.  At N = 8 it is carried as
           a second, fully populated record ``headline_config3`` of the same JSON line.
  config4 : synthetic UHF nbasis=384 (48 alpha, 47 beta), static polarizability.  N = 1: factorised
           AO integrals (the dense tensor is 174 GB); N >= 2: dense tensor sharded over nu.
  config2 : H2O/aug-cc-pVTZ shape (nbasis=92, nocc=5; pyscf is absent, synthetic integrals), 8
           frequencies, ExactInvCholesky; latency bound, wall time is the number.
Inputs are synthetic (seeded), generated on the device outside the timed region.
``check`` compares the property tensors (and M / K samples) of the timed run with the factorised
CPU restatement (oracle/pmr_factorised.py), run once outside the timed region.
"""

from __future__ import annotations

import os
import sys

if "--impl" in sys.argv and "reference" in sys.argv:
    # torchrun exports OMP_NUM_THREADS=1; the reference arm must use every host core
    for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ.pop(_k, None)

import argparse  # noqa: E402
import json  # noqa: E402
import statistics  # noqa: E402
import subprocess  # noqa: E402
import tempfile  # noqa: E402
import time  # noqa: E402
from pathlib import Path  # noqa: E402

import numpy as np  # noqa: E402

REPO = Path(__file__).resolve().parent
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))

METRIC = "cphf_polarizability_fp64_tflops"
UNIT = "TFLOP/s"


# ----------------------------------------------------------------------------------- workloads
def workload(key: str, world: int, args=None):
    if key == "custom":
        n, o = args.n, args.nocc or max(1, args.n // 8)
        nf = args.nfreq or 8
        return {"key": key, "kind": "rhf", "n": n, "o": o, "input": "dense", "solver": "ExactInv",
                "name": f"custom synthetic RHF nbasis={n} nocc={o}, polarizability at {nf} frequencies",
                "pol_freqs": [float(x) for x in np.linspace(0.0, 0.35, nf)] if nf > 1 else [0.0],
                "magnetizability": bool(args.mag), "check_freqs": [0]}
    if key == "config5":
        return {"key": key, "kind": "rhf", "n": 256, "o": 32, "input": "dense", "solver": "ExactInv",
                "name": "config 5: synthetic RHF nbasis=256 nocc=32, magnetizability (static, imaginary) + "
                        "64-frequency polarizability sweep",
                "pol_freqs": [float(x) for x in np.linspace(0.0, 0.63, 64)], "magnetizability": True,
                "check_freqs": [0, 1, 63]}
    if key == "config3":
        return {"key": key, "kind": "rhf", "n": 512, "o": 64, "input": "dense", "solver": "ExactInv",
                "name": "config 3: synthetic RHF nbasis=512 nocc=64, polarizability at 8 frequencies, "
                        "AO ERIs sharded over 8 GPUs",
                "pol_freqs": [float(x) for x in np.linspace(0.0, 0.35, 8)], "magnetizability": False,
                "check_freqs": [0, 7]}
    if key == "config4":
        dense = world > 1
        return {"key": key, "kind": "uhf", "n": 384, "o": (48, 47), "solver": "ExactInv",
                "input": "dense" if dense else "df",
                "name": "config 4: synthetic UHF nbasis=384 (48 alpha, 47 beta), static polarizability, "
                        + ("dense AO ERIs sharded over nu" if dense else
                           "factorised AO integrals (dense tensor = 174 GB does not fit one GPU)"),
                "pol_freqs": [0.0], "magnetizability": False, "check_freqs": [0]}
    if key == "config2":
        return {"key": key, "kind": "rhf", "n": 92, "o": 5, "input": "dense", "solver": "ExactInvCholesky",
                "name": "config 2 shape: synthetic RHF nbasis=92 nocc=5 (H2O/aug-cc-pVTZ dimensions; pyscf "
                        "absent), polarizability at 8 frequencies, ExactInvCholesky",
                "pol_freqs": [float(x) for x in np.linspace(0.0, 0.35, 8)], "magnetizability": False,
                "check_freqs": list(range(8))}
    raise ValueError(key)


def algorithmic_flops(w) -> float:
    from pymolresponse_b200 import ao2mo as ao2mo_mod
    from pymolresponse_b200 import synthetic

    n = w["n"]
    freqs = w["pol_freqs"]
    n_static = sum(1 for f in freqs if f == 0.0)
    n_dyn = len(freqs) - n_static
    if w["kind"] == "uhf":
        na, nb = w["o"]
        va, vb = n - na, n - nb
        nt = na * va + nb * vb
        if w["input"] == "df":
            f = 0.0
            for o, v in ((na, va), (nb, vb)):
                f += 4.0 * n * n**3 + 2.0 * n * (o * v) ** 2 + 2.0 * n * o * o * v * v
            f += 2.0 * 2.0 * n * (na * va) * (nb * vb)
        else:
            f = 2.0 * n**4 * (na + nb)
            for ox, vx in ((na, va), (nb, vb)):     # same-spin (ia|jb) and (ij|ab)
                f += 2 * (2.0 * n**3 * ox * ox + 2.0 * n**2 * ox * ox * vx + 2.0 * n * ox * ox * vx * vx)
            f += 2.0 * n**3 * na * nb + 2.0 * n**2 * na * nb * va + 2.0 * n * na * nb * va * vb
        f += synthetic.flops_solve_rhf(nt, n_static, n_dyn, 3)
        return f
    o = w["o"]
    nov = o * (n - o)
    f = ao2mo_mod.flops_rhf_partial(n, o)
    f += synthetic.flops_solve_rhf(nov, n_static, n_dyn, 3)
    if w["magnetizability"]:
        # static K-solve of the imaginary operator; potrf(K) is shared with the sweep when there
        # is one, otherwise it costs one u
        f += 4.0 * nov * nov * 3
        if n_dyn == 0:
            f += nov**3 / 3.0
    return f


def static_config(w, world):
    """The part of ``config`` both arms print (no measured values: the driver compares them)."""
    n = w["n"]
    if w["kind"] == "uhf":
        na, nb = w["o"]
        nov = [na * (n - na), nb * (n - nb)]
    else:
        nov = w["o"] * (n - w["o"])
    return {"workload": w["name"], "nbasis": n, "nocc": w["o"], "nov": nov,
            "frequencies": len(w["pol_freqs"]), "algorithmic_flops_per_step": algorithmic_flops(w),
            "l2_policy": "inputs larger than L2 (AO ERI tensor / MO blocks in the GBs per GPU)"}


# ----------------------------------------------------------------------------------- helpers
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        try:
            for line in open(self.path):
                t = [x.strip() for x in line.split(",")]
                if len(t) < 9:
                    continue
                try:
                    sm.append(float(t[1])); mx.append(float(t[2])); pw.append(float(t[3]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                      "sw_power_cap"), t[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            hi = [s for s, p in zip(sm, pw) if p >= 0.5 * max(pw)] or sm
            out = {"sm_mhz": statistics.median(hi), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                   "samples": len(sm), "power_w_max": max(pw)}
        return out


def measured_peaks():
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return json.loads(p.read_text()), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback (B200_PROFILING.md)"


def fp64_dgemm_peak(torch):
    """The roofline denominator for the FP64 kernels: cuBLAS DGEMM 8192^3 on this GPU, best of 5
    (burst), measured in this run -- MEASURED_PEAKS.json carries bf16 and HBM only.  cuBLAS is the
    yardstick here, never part of the product path."""
    nn = 8192
    a = torch.randn(nn, nn, dtype=torch.float64, device="cuda")
    b = torch.randn(nn, nn, dtype=torch.float64, device="cuda")
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    torch.cuda.empty_cache()
    return 2.0 * nn**3 / (best * 1e-3) / 1e12


def ncu_traffic(kernel_class: str):
    """Per-launch DRAM bytes of a kernel class from the committed ncu capture of this command
    (profiles/r02_ncu_traffic.json, written by tools/summarize_ncu.py); None when absent."""
    p = REPO / "profiles" / "r02_ncu_traffic.json"
    try:
        return json.loads(p.read_text()).get(kernel_class)
    except Exception:
        return None


def host_inputs(w):
    """Seeded host-side inputs of a workload (everything except the n^4 tensor)."""
    from pymolresponse_b200 import synthetic

    n = w["n"]
    if w["kind"] == "uhf":
        na, nb = w["o"]
        case = synthetic.uhf_case(n, na, nb, build_eri=False)
    else:
        case = synthetic.rhf_case(n, w["o"], build_eri=False)
    return {"C": case["C"], "E": case["E"], "occ": case["occupations"], "B": case["B"],
            "O_real": synthetic.operator_integrals(n, 3, False),
            "O_imag": synthetic.operator_integrals(n, 3, True)}


def build_inputs(w, rank, world, torch):
    """Seeded synthetic inputs; the AO ERI slab I[:, nu_lo:nu_hi] is built on the device from the
    factor B (input generation, cuBLAS, outside the timed region)."""
    from pymolresponse_b200 import _native as nv
    from pymolresponse_b200 import distributed as pd

    inp = host_inputs(w)
    n = w["n"]
    lo, hi = pd.shard_range(n, rank, world)
    inp["nu_range"] = (lo, hi)
    inp["I"] = None
    if w["input"] == "df":
        return inp
    B = inp["B"]
    cnt = hi - lo
    Bd = torch.from_numpy(B).cuda()                          # (naux, n, n)
    Bf = Bd.reshape(B.shape[0], n * n)
    I = torch.empty((n, cnt, n, n), dtype=torch.float64, device="cuda")
    rows_per = max(1, (1 << 28) // (n * n))                  # <= 2 GiB of output per GEMM
    for mu in range(n):
        left = Bd[:, mu, lo:hi]                              # (naux, cnt)
        for r0 in range(0, cnt, rows_per):
            r1 = min(cnt, r0 + rows_per)
            I[mu, r0:r1] = (left[:, r0:r1].t() @ Bf).reshape(r1 - r0, n, n)
    del Bd, Bf
    torch.cuda.empty_cache()
    if world == 1:
        # make the 8-fold permutational symmetry exact to the last bit (cuBLAS rounds (mu nu| and
        # (nu mu| independently), so that the resident and the end-to-end paths see the same tensor
        nv.check(nv.lib().pmr_eri_complete_s8(nv.ptr(I), n, 0, nv.stream_ptr()))
    else:
        nv.check(nv.lib().pmr_eri_complete_ket_s2(nv.ptr(I), n * cnt, n, nv.stream_ptr()))
    torch.cuda.synchronize()
    inp["I"] = I
    return inp


HESSIAN_ROWS = True   # N > 1: row-block-sharded MO blocks / Hessian assembly (--hessian rows)


def one_step(w, inp, I, device_resident: bool, world: int, stage_ms: dict | None = None,
             ao_symmetry: str = "s1", keep_solver: bool = False):
    """One full property calculation through the public API.  ``I`` is a CUDA tensor (kernel
    timing, inputs resident) or a host NumPy array (end-to-end timing, H2D inside)."""
    from pymolresponse_b200 import cphf, integrals, solvers
    from pymolresponse_b200.ao2mo import AO2MO
    from pymolresponse_b200.core import Hamiltonian, Program, Spin
    from pymolresponse_b200.properties.electric import Polarizability
    from pymolresponse_b200.properties.magnetic import Magnetizability

    import torch

    marks = []

    def mark(name):
        if stage_ms is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            marks.append((name, e))

    mark("start")
    nu_range = inp["nu_range"] if world > 1 else None
    if w["input"] == "df":
        a = AO2MO(inp["C"], inp["occ"], df_factor=inp["B"], device_results=True)
    else:
        reduce = False if world == 1 else ("scatter" if HESSIAN_ROWS else True)
        a = AO2MO(inp["C"], inp["occ"], I=I, device_results=True, nu_range=nu_range, reduce=reduce,
                  ao_symmetry=ao_symmetry)
    if w["kind"] == "uhf":
        a.perform_uhf_partial()
    else:
        a.perform_rhf_partial()
    tei = a.tei_mo
    rows = a.row_ranges
    if rows is not None and w["kind"] == "rhf":
        rows = rows[0]
    del a
    mark("ao2mo")
    gen = integrals.IntegralsArrays({"dipole": inp["O_real"], "angmom_common_gauge": inp["O_imag"]})
    out = {}
    s = getattr(solvers, w["solver"])(inp["C"], inp["E"], inp["occ"])
    s.device_results = device_resident
    s.distribute_frequencies = world > 1
    s.tei_mo = tei
    s.tei_mo_rows = rows
    pol = Polarizability(Program.Arrays, gen, cphf.CPHF(s), frequencies=[float(f) for f in w["pol_freqs"]])
    pol.form_operators()
    pol.run(hamiltonian=Hamiltonian.RPA, spin=Spin.singlet)
    pol.form_results()
    out["polarizabilities"] = pol.polarizabilities
    mark("polarizability_sweep")
    if w["magnetizability"]:
        # second property on the same reference: M, K, chol(K) and K^-1 of the sweep are adopted
        # (share_hessian_with), so the static imaginary-operator solve is one product with K^-1
        sm = solvers.ExactInvCholesky(inp["C"], inp["E"], inp["occ"])
        sm.device_results = device_resident
        sm.tei_mo = tei
        sm.tei_mo_rows = rows
        s.share_hessian_with(sm)
        mag = Magnetizability(Program.Arrays, gen, cphf.CPHF(sm))
        mag.form_operators()
        mag.run(hamiltonian=Hamiltonian.RPA, spin=Spin.singlet)
        mag.form_results()
        out["magnetizability"] = mag.magnetizability
        mark("magnetizability")
    out["stats"] = getattr(s, "solve_stats", {})
    out["solver"] = s if keep_solver else None
    if stage_ms is not None:
        torch.cuda.synchronize()
        for (_, e0), (name, e1) in zip(marks[:-1], marks[1:]):
            stage_ms[name] = e0.elapsed_time(e1)
    return out


# ----------------------------------------------------------------------------------- CPU legs
def _all_host_threads():
    """Undo torchrun's OMP_NUM_THREADS=1 for the BLAS behind NumPy/SciPy (reference arm, checks)."""
    cores = os.cpu_count() or 1
    try:
        import threadpoolctl

        threadpoolctl.threadpool_limits(limits=cores)
    except Exception:
        pass
    return cores


# measured seconds per step of the reference algorithm on 16 host cores (oracle port), used only to
# pick the largest sample that keeps the whole --steps/--warmup run within a few minutes
_SAMPLE_COST_S = [(64, 8, 0.2), (96, 12, 1.3), (128, 16, 4.5), (160, 20, 30.0), (192, 24, 90.0)]


def cpu_sample_spec(steps: int, warmup: int, budget_s: float = 240.0):
    per_step = budget_s / max(1, steps + warmup)
    n, o = 64, 8
    for nn, oo, cost in _SAMPLE_COST_S:
        if cost <= per_step:
            n, o = nn, oo
    return {"n": n, "o": o, "freqs": [0.0, 0.21, 0.42, 0.63]}


def cpu_reference_step(spec, case=None, Or=None):
    """The reference's algorithm (oracle port: np.dot + einsum quarter transforms, explicit
    super-Hessian and np.linalg.inv per frequency) on a bounded sample of the workload."""
    from oracle import pmr_oracle as orc
    from pymolresponse_b200 import synthetic

    n, o = spec["n"], spec["o"]
    case = case or synthetic.rhf_case(n, o)
    Or = Or if Or is not None else synthetic.operator_integrals(n, 3, False)
    t0 = time.perf_counter()
    tei = orc.ao2mo_rhf_partial(case["I"], case["C"], o)
    t1 = time.perf_counter()
    out = orc.run_cphf(case["C"], case["E"], case["occupations"], tei, "partial",
                       [{"ao_integrals": Or, "is_imaginary": False}], spec["freqs"])
    t2 = time.perf_counter()
    nov = o * (n - o)
    n_static = sum(1 for f in spec["freqs"] if f == 0.0)
    flops = synthetic.flops_ao2mo_rhf_partial(n, o) + synthetic.flops_solve_rhf(
        nov, n_static, len(spec["freqs"]) - n_static, 3)
    return {"seconds": t2 - t0, "ao2mo_s": t1 - t0, "solve_s": t2 - t1, "flops": flops,
            "results": out["results"]}


def sample_text(spec):
    return (f"synthetic RHF nbasis={spec['n']} nocc={spec['o']}, polarizability at "
            f"{len(spec['freqs'])} frequencies (reference algorithm: einsum AO2MO, explicit (2nov)^2 "
            "Hessian, np.linalg.inv per frequency); TFLOP/s uses the same algorithmic flop model")


def extrapolation_note():
    p = REPO / "profiles" / "r01_cpu_extrapolation.json"
    try:
        d = json.loads(p.read_text())
        return {"label": "EXTRAPOLATED from power-law fits of the reference algorithm's stages "
                         "(tools/cpu_scaling.py); the reference cannot run these sizes on a host",
                "source": "profiles/r01_cpu_extrapolation.json",
                "seconds": {k: v.get("total_s") for k, v in d.get("extrapolated", {}).items()}}
    except Exception:
        return None


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from pymolresponse_b200 import synthetic

    cores = _all_host_threads()
    spec = cpu_sample_spec(args.steps, args.warmup)
    case = synthetic.rhf_case(spec["n"], spec["o"])
    Or = synthetic.operator_integrals(spec["n"], 3, False)
    for _ in range(args.warmup):
        cpu_reference_step(spec, case, Or)
    times, flops = [], 0.0
    for _ in range(args.steps):
        r = cpu_reference_step(spec, case, Or)
        times.append(r["seconds"])
        flops = r["flops"]
    t = sum(times) / len(times)
    val = flops / t / 1e12
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    w = workload(args.workload, world, args)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": static_config(w, world),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample_text(spec), "sample_flops": flops,
                         "blas_threads": _blas_threads(),
                         "full_workload_extrapolation": extrapolation_note()},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def _blas_threads():
    try:
        import threadpoolctl

        return max([int(x.get("num_threads", 0)) for x in threadpoolctl.threadpool_info()] or [0])
    except Exception:
        return None


# ---- check against the factorised CPU restatement (subprocess: keeps a BLAS crash out of the bench)
def check_worker(spec_path: str):
    """``bench.py --check-worker spec.json``: run oracle/pmr_factorised.py on a workload's seeded
    inputs and write the property tensors + M/K probes to ``spec['out']`` (npz)."""
    spec = json.loads(Path(spec_path).read_text())
    _all_host_threads()
    from oracle import pmr_factorised as fac

    w = spec["workload"]
    inp = host_inputs(w)
    freqs = [w["pol_freqs"][i] for i in w["check_freqs"]]
    ops = [{"ao_integrals": inp["O_real"], "is_imaginary": False}]
    if w["magnetizability"]:
        ops.append({"ao_integrals": inp["O_imag"], "is_imaginary": True})
    t0 = time.perf_counter()
    out = fac.run_cphf_factorised(inp["B"], inp["C"], inp["E"], inp["occ"], ops, freqs,
                                  return_mk_samples=spec.get("mk_samples", CHECK_MK_SAMPLES),
                                  seed=spec.get("mk_seed", CHECK_MK_SEED))
    seconds = time.perf_counter() - t0
    smp = out["mk_samples"]
    np.savez(spec["out"], results=np.stack(out["results"]), seconds=seconds,
             rows=np.array([s[0] for s in smp[:-1]], dtype=np.int64),
             cols=np.array([s[1] for s in smp[:-1]], dtype=np.int64),
             M=np.array([s[2] for s in smp[:-1]]), K=np.array([s[3] for s in smp[:-1]]),
             M_max=smp[-1][2], K_max=smp[-1][3])


_T0 = time.time()
CHECK_MK_SAMPLES, CHECK_MK_SEED = 64, 7
WALL_BUDGET_S = 780.0     # the driver gives every bench invocation 870 s


def run_check(w, gpu, timeout_s=None):
    """gpu: dict(polarizabilities: {freq index: (3,3)}, magnetizability or None, mk: callable
    (rows, cols) -> (M values, K values) or None).  Returns the ``check`` record."""
    tmp = tempfile.mkdtemp(prefix="pmr_check_")
    spec = {"workload": w, "out": os.path.join(tmp, "cpu.npz"), "mk_samples": CHECK_MK_SAMPLES,
            "mk_seed": CHECK_MK_SEED}
    sp = os.path.join(tmp, "spec.json")
    Path(sp).write_text(json.dumps(spec))
    env = dict(os.environ)
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        env.pop(k, None)
    env["CUDA_VISIBLE_DEVICES"] = ""
    rec = {"against": "oracle/pmr_factorised.py (factorised CPU restatement, LAPACK Cholesky)",
           "frequencies_checked": [w["pol_freqs"][i] for i in w["check_freqs"]],
           "tolerance_property_tensors": 1e-8, "tolerance_mk_rel": 1e-10}
    if timeout_s is None:
        timeout_s = max(30.0, WALL_BUDGET_S - (time.time() - _T0))
    try:
        t0 = time.perf_counter()
        r = subprocess.run([sys.executable, str(REPO / "bench.py"), "--check-worker", sp],
                           env=env, timeout=timeout_s, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        rec["cpu_seconds"] = time.perf_counter() - t0
        if r.returncode != 0:
            rec["error"] = f"check worker exited {r.returncode}: {r.stderr.decode(errors='replace')[-300:]}"
            return rec
        cpu = np.load(spec["out"])
    except subprocess.TimeoutExpired:
        rec["error"] = (f"check worker stopped after {timeout_s:.0f} s (wall budget of the bench run); "
                        "run `python bench.py --workload ... --steps 1` for the check alone")
        return rec
    worst = 0.0
    for k, idx in enumerate(w["check_freqs"]):
        ref = cpu["results"][k]
        worst = max(worst, float(np.abs(np.asarray(gpu["polarizabilities"][idx]) - ref[:3, :3]).max()))
        if k == 0 and w["magnetizability"] and gpu.get("magnetizability") is not None:
            dm = float(np.abs(np.asarray(gpu["magnetizability"]) - ref[3:, 3:] / 4.0).max())
            rec["max_abs_err_magnetizability"] = dm
            worst = max(worst, dm)
    rec["max_abs_err_vs_cpu_restatement"] = worst
    rec["alpha_static_diag"] = [float(x) for x in np.diag(np.asarray(gpu["polarizabilities"][0]))]
    if gpu.get("mk") is not None:
        rows_i, cols_i, Mg, Kg = gpu["mk"]
        assert np.array_equal(rows_i, cpu["rows"]) and np.array_equal(cols_i, cpu["cols"])
        rec["max_rel_err_M_samples"] = float(np.abs(Mg - cpu["M"]).max() / float(cpu["M_max"]))
        rec["max_rel_err_K_samples"] = float(np.abs(Kg - cpu["K"]).max() / float(cpu["K_max"]))
        rec["mk_samples"] = int(len(cpu["rows"]))
    rec["pass"] = bool(worst <= 1e-8 and rec.get("max_rel_err_M_samples", 0.0) <= 1e-10
                       and rec.get("max_rel_err_K_samples", 0.0) <= 1e-10)
    return rec


# ----------------------------------------------------------------------------------- ours
def bench_workload(w, args, rank, world, local, torch, peak64, do_e2e=True, do_clocks=True,
                   e2e_steps_cap=0):
    """Time one workload: resident steps (value), roofline pass, stage clocks, e2e variants, check."""
    from pymolresponse_b200 import _engine as eng
    from pymolresponse_b200 import _native as nv
    from pymolresponse_b200 import ao2mo as ao2mo_mod
    from pymolresponse_b200 import distributed as pd

    lib = nv.lib()
    F = algorithmic_flops(w)
    note(f"{w['key']}: building inputs")
    inp = build_inputs(w, rank, world, torch)
    I_dev = inp["I"]
    note(f"{w['key']}: inputs resident, warm-up + timed steps")

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            pd.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            pd.dist().all_reduce(t, op=pd.dist().ReduceOp.MAX)
        return float(t.item())

    # ---------------- kernel-only timing: inputs resident in HBM -----------------------------
    for _ in range(args.warmup):
        res = one_step(w, inp, I_dev, True, world)
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0 and do_clocks:
        sampler.start()
    nv.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        res = None       # the previous step's matrices are released first, as a caller's loop would
        res = one_step(w, inp, I_dev, True, world, keep_solver=True)
    e1.record()
    sync_all()
    launches = nv.launch_count()
    clocks = sampler.stop() if (rank == 0 and do_clocks) else None
    ms = max_over_ranks(e0.elapsed_time(e1) / args.steps)

    # results of the timed run, for the check (small D2H, outside the timed region); the M / K probes
    # are read now so that the solver's matrices can be released before the end-to-end variants
    gpu = {"polarizabilities": {i: nv.to_host(res["polarizabilities"][i]) for i in w["check_freqs"]},
           "magnetizability": nv.to_host(res["magnetizability"]) if w["magnetizability"] else None,
           "mk": None}
    system = getattr(res["solver"], "_system", None)
    if (system is not None and getattr(system, "M", None) is not None
            and int(system.M.shape[0]) == int(system.N) and not args.no_check):
        from oracle import pmr_factorised as fac   # check leg only: the probe positions

        rows_i, cols_i = fac.sample_indices(int(system.N), CHECK_MK_SAMPLES, CHECK_MK_SEED)
        rt, ct = torch.from_numpy(rows_i).cuda(), torch.from_numpy(cols_i).cuda()
        gpu["mk"] = (rows_i, cols_i, system.M[rt, ct].cpu().numpy(), system.K[rt, ct].cpu().numpy())
        del rt, ct
    del system
    res = {"polarizabilities": [nv.to_host(p_) for p_ in res["polarizabilities"]],
           "stats": res.get("stats")}
    # (no empty_cache here: the stage-clock pass below must not pay for fresh cudaMallocs)

    # ---------------- stage clocks + roofline pass (untimed extra steps) ---------------------
    note(f"{w['key']}: resident {ms:.1f} ms/step; stage clocks + roofline pass")
    stages = {}
    eng.STAGE_TIMES = True
    ao2mo_mod.LAST_STAGE.clear()
    sync_all()      # rank 0 has just spent 0.2 s stopping the clock sampler: start the pass together
    res_s = one_step(w, inp, I_dev, True, world, stage_ms=stages)
    eng.STAGE_TIMES = False
    stages["ao2mo_parts"] = dict(ao2mo_mod.LAST_STAGE.get("ms", {}))
    stages["sweep_parts"] = dict((res_s.get("stats") or {}).get("ms", {}))
    del res_s
    sync_all()

    roofline = None
    if rank == 0:
        lib.pmr_profile_enable(1)
        lib.pmr_set_lookahead(0)   # serialise the Cholesky so every launch is timed on its own
    one_step(w, inp, I_dev, True, world)   # every rank takes part (collectives), rank 0 records
    sync_all()
    if rank == 0:
        lib.pmr_set_lookahead(1)
        import ctypes as C

        buf = (C.c_double * 24)()
        lib.pmr_profile_fetch(buf, 6)
        lib.pmr_profile_enable(0)
        prof = {}
        names = ["gemm128x128", "gemm128x128_lower(trailing update)", "gemm128x64", "gemm128x64_lower",
                 "gemm128x32", "gemm128x32_lower"]
        for c in range(6):
            cnt, tms, fl, by = buf[c * 4], buf[c * 4 + 1], buf[c * 4 + 2], buf[c * 4 + 3]
            if cnt > 0:
                prof[names[c]] = {"launches": int(cnt), "ms": tms, "tflops": fl / (tms * 1e-3) / 1e12,
                                  "algorithmic_flops": fl, "algorithmic_bytes": by}
        if prof:
            top = max(prof, key=lambda k: prof[k]["ms"])
            a = prof[top]
            tr = ncu_traffic(f"{w['key']}:n{world}:{top}")
            roofline = {"bound": "tensor", "kernel": f"dgemm kernel class [{top}]",
                        "achieved": a["tflops"], "peak": peak64, "unit": "TFLOP/s",
                        "frac": a["tflops"] / peak64 if peak64 else None,
                        "traffic": tr.get("dram_bytes_per_launch") if tr else None,
                        "traffic_source": tr.get("source") if tr else None,
                        "algorithmic_flops_per_launch": a["algorithmic_flops"] / a["launches"],
                        "algorithmic_bytes_per_launch": a["algorithmic_bytes"] / a["launches"],
                        "peak_source": "cuBLAS DGEMM 8192^3 (torch.matmul fp64) measured in this run, "
                                       "burst; MEASURED_PEAKS.json has no FP64 figure; nominal B200 "
                                       "FP64 is 37-40 TFLOP/s",
                        "share_of_step": a["ms"] / ms, "avg_launch_ms": a["ms"] / a["launches"],
                        "classes": prof}

    # ---------------- end-to-end: host buffers, H2D + D2H inside the timed region ------------
    e2e = None
    note(f"{w['key']}: end-to-end variants")
    if do_e2e and not args.no_e2e and w["input"] == "dense":
        e2e = bench_e2e(w, args, inp, res, rank, world, torch, F, sync_all, max_over_ranks,
                        e2e_steps_cap)
    elif do_e2e and not args.no_e2e:
        # factorised input: the factor B (naux n^2 doubles) is the host input of every step
        small = sum(inp[k].nbytes for k in ("C", "E", "O_real", "O_imag", "B"))
        one_step(w, inp, None, False, world)
        sync_all()
        e0.record()
        for _ in range(args.steps):
            r = one_step(w, inp, None, False, world)
        e1.record()
        sync_all()
        ms_e = max_over_ranks(e0.elapsed_time(e1) / args.steps)
        nt = sum(static_config(w, world)["nov"])
        e2e = {"value": F / (ms_e * 1e-3) / 1e12, "unit": UNIT, "ms_per_step": ms_e,
               "h2d_bytes_per_step": int(small), "d2h_bytes_per_step": int(3 * 2 * nt * 8 + 72),
               "upload": "factor B (pageable NumPy) + C, E, operators; response vectors and results "
                         "read back as NumPy", "steps": args.steps}
        del r
    inp["I"] = None
    del I_dev

    record = {
        "value": F / (ms * 1e-3) / 1e12, "ms_per_step": ms, "steps": args.steps, "warmup": args.warmup,
        "config": static_config(w, world), "clocks": clocks, "e2e": e2e,
        "gpu_launches": int(launches // max(1, args.steps)), "roofline": roofline,
        "stage_ms": stages, "solve_stats": {k: v for k, v in (res.get("stats") or {}).items() if k != "ms"},
        "fraction_of_fp64_peak": (F / (ms * 1e-3) / 1e12) / (world * peak64) if peak64 else None,
    }
    del res
    torch.cuda.empty_cache()
    return record, gpu


def note(msg):
    """Progress line on stderr (rank 0): which phase a killed run died in, with the host memory left."""
    if int(os.environ.get("RANK", "0")) == 0:
        a = host_memory_available()
        sys.stderr.write(f"[bench {time.time() - _T0:7.1f}s] {msg} (host memory left: "
                         f"{'?' if a is None else f'{a / 2**30:.0f} GiB'})\n")
        sys.stderr.flush()


def host_memory_available():
    """Bytes of host memory this process may still take: the smaller of what the machine reports and
    what the container's cgroup (v2 memory.max / v1 memory.limit_in_bytes) leaves -- a pinned
    allocation beyond the cgroup limit gets the whole job killed, not an exception."""
    avail = None
    try:
        import psutil

        avail = int(psutil.virtual_memory().available)
    except Exception:
        pass
    for lim_p, use_p in (("/sys/fs/cgroup/memory.max", "/sys/fs/cgroup/memory.current"),
                         ("/sys/fs/cgroup/memory/memory.limit_in_bytes",
                          "/sys/fs/cgroup/memory/memory.usage_in_bytes")):
        try:
            lim = open(lim_p).read().strip()
            if lim == "max":
                continue
            left = int(lim) - int(open(use_p).read().strip())
            if 0 < int(lim) < (1 << 60):
                avail = left if avail is None else min(avail, left)
        except Exception:
            continue
    return avail


def bench_e2e(w, args, inp, res, rank, world, torch, F, sync_all, max_over_ranks, steps_cap=0):
    """The same step through the public API with the AO tensor in HOST memory (H2D inside the timed
    region, response vectors + property tensors read back as NumPy).  Variants:
      s1_pinned   -- whole array, pinned host buffer            (`e2e` itself: the reference API, no
                     assumption about the tensor; pinned as the bench contract prescribes)
      s1_pageable -- whole array, plain pageable NumPy array    (what a drop-in caller hands over)
      s8_pinned   -- ao_symmetry='s8' opt-in: permutationally unique part only, completed in HBM"""
    from pymolresponse_b200 import _native as nv
    from pymolresponse_b200 import distributed as pd

    lib = nv.lib()
    I_dev = inp["I"]
    n = w["n"]
    need = I_dev.numel() * 8
    avail = host_memory_available()
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    # pinned copy of every rank's slab (level 1) plus a pageable copy (level 2), with head room: the
    # staging buffers, the CPU check and the interpreter itself live in the same cgroup
    # (measured: 512 GiB pinned of 954 GiB available ran, 162 GiB pinned of 248 GiB got the job
    # OOM-killed -- pinned pages, CUDA host contexts and the staging buffers add up)
    level = 0 if avail is None else (2 if avail > 2.9 * need * local_world + (8 << 30) else
                                     1 if avail > 1.75 * need * local_world + (8 << 30) else -1)
    if avail is None:
        level = 1
    flag = torch.tensor([level], dtype=torch.int64, device="cuda")
    if world > 1:
        pd.dist().all_reduce(flag, op=pd.dist().ReduceOp.MIN)
    level = int(flag.item())
    note(f"e2e: need {need / 2**30:.0f} GiB per rank x {local_world} ranks, level {level}")
    if level < 0:
        return {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "note": f"skipped: host RAM available {avail / 2**30:.0f} GiB < {local_world} ranks x "
                        f"{need / 2**30:.0f} GiB AO-tensor slabs"}
    try:
        I_pin_t = torch.empty(I_dev.shape, dtype=torch.float64, pin_memory=True)
    except Exception:
        I_pin_t = torch.empty(I_dev.shape, dtype=torch.float64)
    I_pin_t.copy_(I_dev)
    torch.cuda.synchronize()
    I_pin = I_pin_t.numpy()
    inp["I"] = None
    del I_dev
    torch.cuda.empty_cache()
    steps = args.steps if args.e2e_steps <= 0 else max(1, min(args.steps, args.e2e_steps))
    if steps_cap:
        steps = max(1, min(steps, steps_cap))
    small = sum(inp[k].nbytes for k in ("C", "E", "O_real", "O_imag"))
    nov = static_config(w, world)["nov"]
    nt = sum(nov) if isinstance(nov, list) else nov
    nfreq = len(w["pol_freqs"])
    d2h = nfreq * (3 * 2 * nt * 8 + 72) + (3 * 2 * nt * 8 + 72 if w["magnetizability"] else 0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    diffs = []

    def timed(I_host, sym, nsteps=None):
        nsteps = steps if nsteps is None else nsteps
        one_step(w, inp, I_host, False, world, ao_symmetry=sym)
        sync_all()
        e0.record()
        for _ in range(nsteps):
            r = one_step(w, inp, I_host, False, world, ao_symmetry=sym)
        e1.record()
        sync_all()
        ms_e = max_over_ranks(e0.elapsed_time(e1) / nsteps)
        diff = max([float(np.abs(np.asarray(a) - nv.to_host(b)).max())
                    for a, b in zip(r["polarizabilities"], res["polarizabilities"])] or [0.0])
        assert diff <= 1e-8, f"end-to-end and resident results differ by {diff}"
        diffs.append(diff)
        return ms_e

    def rec(ms_e, h2d, pinned, upload, nsteps=None):
        return {"value": F / (ms_e * 1e-3) / 1e12, "unit": UNIT, "ms_per_step": ms_e,
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "pinned_host_input": bool(pinned), "upload": upload,
                "steps": steps if nsteps is None else nsteps}

    whole = int(I_pin.nbytes * world + small)
    variants = {}
    e2e = rec(timed(I_pin, "s1"), whole, I_pin_t.is_pinned(),
              "whole AO ERI array (ao_symmetry='s1', the reference API; nothing assumed)")
    if world == 1:
        s8_bytes = int(lib.pmr_eri_s8_bytes(n, nv.ERI_S8_LAYOUT)) + int(small)
        s8_text = ("permutationally unique part of the AO ERI array (opt-in ao_symmetry='s8': rows "
                   "(la, si <= la), columns from (la, 0) on -- a sixth of the tensor in long runs), "
                   "completed on the device")
    else:
        s8_bytes = int(world * lib.pmr_eri_ket_s2_bytes(I_pin.shape[0] * I_pin.shape[1], n)) + int(small)
        s8_text = ("ket-symmetric half (si <= la) of every rank's nu-slab (opt-in ao_symmetry='s8'), "
                   "completed on the device")
    variants["s8_pinned"] = rec(timed(I_pin, "s8"), s8_bytes, I_pin_t.is_pinned(), s8_text)
    if not args.no_pageable and level >= 2:
        I_page = np.empty(I_pin.shape, dtype=np.float64)     # plain pageable NumPy array
        np.copyto(I_page, I_pin)
        npg = max(1, min(steps, 5))     # seconds per step: a few steps are enough
        variants["s1_pageable"] = rec(timed(I_page, "s1", npg), whole, False,
                                      "whole AO ERI array from a plain pageable NumPy array "
                                      "(ao_symmetry='s1'): the literal drop-in call", npg)
        del I_page
    e2e["variants"] = variants
    e2e["max_abs_diff_vs_resident"] = max(diffs) if diffs else None
    del I_pin_t, I_pin
    return e2e


def same_config_sample(args, torch, world):
    """The GPU path timed on EXACTLY the reference arm's bounded sample (one GPU, public API):
    resident and end-to-end (plain NumPy AO tensor in, NumPy results out), checked against the
    reference algorithm's results for that sample."""
    from pymolresponse_b200 import cphf, operators, solvers, synthetic
    from pymolresponse_b200.core import Hamiltonian, Program, Spin

    spec = cpu_sample_spec(args.steps, args.warmup)
    n, o = spec["n"], spec["o"]
    case = synthetic.rhf_case(n, o)
    Or = synthetic.operator_integrals(n, 3, False)
    I_dev = torch.from_numpy(case["I"]).cuda()

    def step(I, resident):
        s = solvers.ExactInv(case["C"], case["E"], case["occupations"])
        s.device_results = resident
        d = cphf.CPHF(s)
        d.add_operator(operators.Operator(label="dipole", is_imaginary=False, ao_integrals=Or))
        d.set_frequencies(spec["freqs"])
        d.run(hamiltonian=Hamiltonian.RPA, spin=Spin.singlet, program=Program.Arrays, program_obj=I)
        return d.results

    out = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for tag, I, resident in (("resident", I_dev, True), ("e2e", case["I"], False)):
        for _ in range(max(3, args.warmup)):
            step(I, resident)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.steps):
            results = step(I, resident)
        e1.record()
        torch.cuda.synchronize()
        out[tag] = e0.elapsed_time(e1) / args.steps
    from pymolresponse_b200 import _native as nv

    cores = _all_host_threads()
    ref = cpu_reference_step(spec, case, Or)
    err = max(float(np.abs(nv.to_host(results[f]) - ref["results"][f]).max())
              for f in range(len(spec["freqs"])))
    return {
        "sample": sample_text(spec), "sample_flops": ref["flops"],
        "gpu_resident_ms_per_step": out["resident"], "gpu_e2e_ms_per_step": out["e2e"],
        "gpu_value": ref["flops"] / (out["resident"] * 1e-3) / 1e12,
        "gpu_e2e_value": ref["flops"] / (out["e2e"] * 1e-3) / 1e12, "unit": UNIT,
        "h2d_bytes_per_step": int(case["I"].nbytes),
        "cpu_reference_seconds_per_step": ref["seconds"], "cpu_cores": cores,
        "cpu_value": ref["flops"] / ref["seconds"] / 1e12,
        "max_abs_err_vs_reference_algorithm": err,
        "note": "same workload on both sides: the reference arm's bounded sample, one GPU",
    }, {"value": ref["flops"] / ref["seconds"] / 1e12, "unit": UNIT, "cores": cores, "kind": "port",
        "sample": sample_text(spec), "seconds": ref["seconds"], "blas_threads": _blas_threads(),
        "full_workload_extrapolation": extrapolation_note()}


def run_ours(args):
    import torch

    from pymolresponse_b200 import distributed as pd

    rank, world, local = pd.init_from_env()
    torch.cuda.set_device(local)
    w = workload(args.workload, world, args)
    peak64 = fp64_dgemm_peak(torch)
    if world > 1:
        t = torch.tensor([peak64], dtype=torch.float64, device="cuda")
        pd.dist().broadcast(t, src=0)
        peak64 = float(t.item())

    record, gpu = bench_workload(w, args, rank, world, local, torch, peak64)
    headline = gpu3 = w3 = None
    if world >= 8 and args.workload == "config5" and not args.no_config3:
        w3 = workload("config3", world, args)
        headline, gpu3 = bench_workload(w3, args, rank, world, local, torch, peak64,
                                        do_e2e=not args.no_config3_e2e, do_clocks=True, e2e_steps_cap=2)

    if world > 1:
        pd.barrier()
        pd.dist().destroy_process_group()
    if rank != 0:
        return
    # rank 0 alone from here on (the other ranks have exited: the host cores are free for the CPU legs)
    sample = cpu = None
    if not args.no_cpu:
        sample, cpu = same_config_sample(args, torch, world)
    # the CPU restatement runs after the process group is gone (minutes at config 3: no collective
    # may be waiting on rank 0 meanwhile)
    note("CPU restatement check")
    check = None if args.no_check else run_check(w, gpu)
    if headline is not None:
        headline["check"] = None if args.no_check else run_check(w3, gpu3)
        headline["metric"], headline["unit"] = METRIC, UNIT
        headline["scaling"] = "n/a (single point: the workload needs 8 GPUs)"
    peaks, src = measured_peaks()
    line = {
        "metric": METRIC, "value": record["value"], "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": record["ms_per_step"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": record["config"],
        "peaks": {"fp64_dgemm_tflops": peak64, "hbm_gbs": peaks.get("hbm_gbs"), "source": src,
                  "fraction_of_fp64_peak": record["fraction_of_fp64_peak"]},
        "clocks": record["clocks"], "e2e": record["e2e"], "gpu_launches": record["gpu_launches"],
        "roofline": record["roofline"], "cpu_baseline": cpu, "same_config_sample": sample,
        "stage_ms": record["stage_ms"], "solve_stats": record["solve_stats"], "check": check,
        "sharding": None if world == 1 else {
            "ao_eri": "second index of the first AO pair (nu)", "frequencies": "round robin over ranks",
            "orbital_hessian": "row blocks by occupied index, all-gathered" if HESSIAN_ROWS
            else "replicated (blocks all-reduced)"},
    }
    if headline is not None:
        line["headline_config3"] = headline
    emit(line)


def _claim_stdout():
    """Route everything libraries print on fd 1 (NCCL banners ...) to stderr and keep a private
    handle on the real stdout, so rank 0 prints exactly ONE JSON line there."""
    real = os.fdopen(os.dup(1), "w")
    sys.stdout.flush()
    os.dup2(2, 1)
    return real


_REAL_STDOUT = None


def emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    global _REAL_STDOUT
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config5",
                    choices=["config5", "config3", "config4", "config2", "custom"])
    ap.add_argument("--n", type=int, default=0, help="nbasis of --workload custom")
    ap.add_argument("--nocc", type=int, default=0)
    ap.add_argument("--nfreq", type=int, default=0)
    ap.add_argument("--mag", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-pageable", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--no-config3", action="store_true", help="N = 8: skip the headline_config3 record")
    ap.add_argument("--no-config3-e2e", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0: as many as --steps")
    ap.add_argument("--hessian", default="rows", choices=["rows", "replicated"],
                    help="N > 1: assemble the orbital Hessian from row-block-sharded MO blocks (rows, "
                         "all-gathered) or on every rank from all-reduced blocks (replicated)")
    ap.add_argument("--check-worker", default=None, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.check_worker:
        check_worker(args.check_worker)
        return
    if args.n and args.workload == "config5":
        args.workload = "custom"
    global HESSIAN_ROWS
    HESSIAN_ROWS = args.hessian == "rows"
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    _REAL_STDOUT = _claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
