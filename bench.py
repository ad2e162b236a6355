#!/usr/bin/env python
"""bench.py -- elements assembled/s of the hot path (BASELINE.json metric) on N B200s of one node.

A "step" = one pass of the numeric assembly (element quadrature loop fused with the deterministic
scatter into nzval) over the whole workload, with the mesh / scatter maps already resident in HBM.
`e2e` = the same metric through the reference-facing C-ABI call sequence with HOST buffers
(efg_set_mesh/_space -> efg_start -> efg_assemble (symbolic + numeric) -> efg_fetch_csc), i.e.
host->device and device->host copies and the symbolic phase inside the timed region.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload heat_t6|heat_q4|heat_t3|elasticity_t6|stokes_gen]
                  [--n SIZE] [--impl reference]

N > 1: launched by torchrun, one rank per GPU; the matrix columns are split into N contiguous blocks
(owner computes, halo elements replicated, no data-path collective); NCCL only carries the barrier,
the max-over-ranks time and validation checksums.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (default N, description)
    "heat_t6": (4000, "heat Poisson FEH1_T6, unit-square NxN T6block (BASELINE config 2)"),
    "heat_t3": (100, "heat Poisson FEH1_T3, unit-square NxN T3block (BASELINE config 1)"),
    "heat_q4": (5792, "heat Poisson FEH1_Q4 Gauss order 2, NxN Q4block (BASELINE config 5 per-GPU share)"),
    "elasticity_t6": (2000, "plane-stress elasticity vector FEH1_T6, NxN T6block (BASELINE config 3)"),
    "stokes_gen": (1000, "Stokes Taylor-Hood T6/T3 'gen' 3-block assembly (BASELINE config 4)"),
}


def make_problem(efg, workload, n):
    if workload == "heat_t6":
        return efg.heat_problem(efg.T6, n)
    if workload == "heat_t3":
        return efg.heat_problem(efg.T3, n)
    if workload == "heat_q4":
        return efg.heat_problem(efg.Q4, n)
    if workload == "elasticity_t6":
        return efg.elasticity_problem(n, efg.T6)
    if workload == "stokes_gen":
        return efg.stokes_problem(n, "gen")
    raise SystemExit(f"unknown workload {workload}")


def triplets_per_element(prob):
    nd = {1: 1, 2: 4}
    k = prob.meshes[0].kind
    fid = prob.form.form_id
    if fid == 1:
        return k * k
    if fid == 2:
        return 4 * k * k
    if fid in (3, 5, 4):
        return 216
    return 144


def algorithmic_bytes(prob, nnz):
    """every input read once + slot map once + every output value written once, int32 device
    indices: 4*nen*nel + 16*nnodes + 4*ndofs + 4*ntriplets + 8*nnz (pressure mesh connectivity added for Stokes)."""
    b = 0
    for m in prob.meshes:
        b += 4 * m.kind * m.nel_
    b += 16 * prob.meshes[0].nnodes_
    b += 4 * getattr(prob, "ndofs_local_", prob.ndofs)     # dof map entries of this rank's sub-mesh
    b += 4 * triplets_per_element(prob) * prob.meshes[0].nel_
    b += 8 * nnz
    return b


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def result(self):
        self.stop_flag = True
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_baseline(efg, workload, sample_n):
    """The oracle (C restatement of the reference's serial COO + sparse() path) on one host core."""
    from oracle import oracle as orc
    orc.build()
    prob = make_problem(efg, workload, sample_n)
    tm = {}
    t0 = time.perf_counter()
    _, _, nz = orc.assemble(*efg.oracle_args(prob), prob.ndofs, prob.ndofs, timing=tm)
    dt = time.perf_counter() - t0
    return {"value": prob.nel / dt, "unit": "elements/s", "cores": 1, "kind": "port",
            "sample": f"{workload} N={sample_n}: {prob.nel} elements, integrate {tm['integrate_s']:.2f} s + finish(sparse) "
                      f"{tm['finish_s']:.2f} s; C restatement of Elfel's CPU path (gcc -O2, no FMA), not Julia"}


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm (oracle port; Julia is absent) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import elfel_jl_b200 as efg
    from oracle import oracle as orc
    orc.build()
    n = args.ref_n
    prob = make_problem(efg, args.workload, n)
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        orc.assemble(*efg.oracle_args(prob), prob.ndofs, prob.ndofs)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    v = prob.nel / (ms / 1e3)
    sample = f"{args.workload} N={n} ({prob.nel} elements) per step; C restatement of Elfel's serial COO + sparse() path, 1 thread"
    print(json.dumps({
        "impl": "reference", "metric": "elements assembled/s", "value": v, "unit": "elements/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{WORKLOADS[args.workload][1]}, bounded sample N={n}"},
        "cpu_baseline": {"value": v, "unit": "elements/s", "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "elements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="heat_t6", choices=sorted(WORKLOADS))
    ap.add_argument("--n", type=int, default=0, help="mesh subdivisions per side (0 = the BASELINE size)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-n", type=int, default=500)
    ap.add_argument("--cpu-n", type=int, default=1000)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-warmup", type=int, default=4, help="untimed e2e calls (the device memory pool needs a few calls to reach steady state)")
    ap.add_argument("--tile-elems", type=int, default=0)
    ap.add_argument("--sfc", type=int, default=1)
    ap.add_argument("--strict", type=int, default=0)
    ap.add_argument("--path", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-callers", action="store_true", help="skip the f1/f2 rows (load vector, K*x) timed after the hot path")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import elfel_jl_b200 as efg
    from elfel_jl_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the assembly path")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    n = args.n or WORKLOADS[args.workload][0]
    if world > 1:
        # weak scaling: a mesh `world` units tall, rank r owns the dofs of horizontal band r (a set of column
        # ranges), halo elements replicated; generated on the GPU (elfel.jl_b200/sharding.py)
        from elfel_jl_b200.sharding import shard_problem
        prob, col_range, nel_global = shard_problem(efg, args.workload, n, rank, world)
    else:
        prob, col_range, nel_global = make_problem(efg, args.workload, n), None, None
    nel_global = nel_global or prob.nel

    eng = efg.Engine(local)
    eng.set_option(_lib.OPT_PATH, args.path)
    eng.set_option(_lib.OPT_STRICT_FP, args.strict)
    eng.set_option(_lib.OPT_TILE_ELEMS, args.tile_elems)
    eng.set_option(_lib.OPT_SFC_ORDER, args.sfc)
    stream = torch.cuda.ExternalStream(eng.stream(), device=torch.device("cuda", local))
    params = prob.form.params()

    # host buffers (pinned) of the reference-facing call
    def pin(a):
        t = a.detach().cpu().contiguous() if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
        return t.pin_memory()
    h_mesh = [(m.kind, pin(m.conn), pin(m.xy)) for m in prob.meshes]
    h_dofs = [pin(s.field.dofnums) for s in prob.spaces]
    for m in prob.meshes:           # the host copies are the inputs from here on
        m.conn_shape, m.nnodes_, m.nel_ = tuple(m.conn.shape), int(m.xy.shape[0]), int(m.conn.shape[0])
    if world > 1:
        for m in prob.meshes:
            m.conn = m.xy = None
        for sp in prob.spaces:
            sp.field.dofnums = None
        torch.cuda.empty_cache()
    h2d = sum(c.numel() * 8 + x.numel() * 8 for _, c, x in h_mesh) + sum(d.numel() * 8 for d in h_dofs)
    prob.ndofs_local_ = int(sum(d.numel() for d in h_dofs))

    def load():
        for slot, (kind, c, x) in enumerate(h_mesh):
            eng.set_mesh(slot, kind, c, x)
        for slot, (d, ms) in enumerate(zip(h_dofs, prob.space_mesh)):
            eng.set_space(slot, ms, d)
        eng.start(prob.ndofs, prob.ndofs)
        if col_range is not None:
            eng.set_column_ranges(*col_range)

    # ---- e2e through the C ABI with host buffers -------------------------------------------------------
    e2e = None
    load()
    nnz = eng.assemble(prob.form.form_id, prob.quad, params)
    eng.synchronize()
    sym_ms = eng.stat(_lib.STAT_SYMBOLIC_MS)
    if not args.no_e2e:
        ncl = eng.ncols_local
        o_colptr = torch.empty(ncl + 1, dtype=torch.int64).pin_memory()
        o_rowval = torch.empty(nnz, dtype=torch.int64).pin_memory()
        o_nzval = torch.empty(nnz, dtype=torch.float64).pin_memory()
        d2h = 8 * (ncl + 1) + 16 * nnz
        ts, all_ts, sym_hist = [], [], []
        for it in range(args.e2e_warmup + args.e2e_steps):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            load()
            eng.assemble(prob.form.form_id, prob.quad, params)
            eng.fetch_csc(o_colptr, o_rowval, o_nzval)
            eng.synchronize()
            dt = time.perf_counter() - t0
            if world > 1:
                tt = torch.tensor([dt], device="cuda")
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                dt = float(tt.item())
            all_ts.append(round(1e3 * dt, 1))
            sym_hist.append(round(eng.stat(_lib.STAT_SYMBOLIC_MS), 1))
            if it >= args.e2e_warmup:
                ts.append(dt)
        e2e_s = float(np.mean(ts))
        e2e = {"value": nel_global / e2e_s, "unit": "elements/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e2e_s, "steps": args.e2e_steps, "warmup": args.e2e_warmup,
               "all_calls_ms": all_ts, "symbolic_ms_per_call": sym_hist,
               "what": "efg_set_mesh/_space + efg_start + efg_assemble (symbolic+numeric) + efg_fetch_csc, pinned host buffers"}
        checksum = float(o_nzval.sum().item())
        del o_colptr, o_rowval, o_nzval
    else:
        checksum = None

    # ---- device-resident numeric phase (the hot path), CUDA events on the library's stream --------------
    launches0 = eng.stat(_lib.STAT_KERNEL_LAUNCHES)
    for _ in range(args.warmup):
        eng.numeric(params)
    eng.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches1 = eng.stat(_lib.STAT_KERNEL_LAUNCHES)
    ev0.record(stream)
    for _ in range(args.steps):
        eng.numeric(params)
    ev1.record(stream)
    ev1.synchronize()
    torch.cuda.synchronize()
    clocks = sampler.result()
    gpu_launches = int(eng.stat(_lib.STAT_KERNEL_LAUNCHES) - launches1)
    ms_total = ev0.elapsed_time(ev1)
    if world > 1:
        tt = torch.tensor([ms_total], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt.item())
        tn = torch.tensor([float(nnz)], device="cuda", dtype=torch.float64)
        dist.all_reduce(tn)
        nnz_global = int(tn.item())
    else:
        nnz_global = nnz
    ms_step = ms_total / args.steps
    value = nel_global / (ms_step / 1e3)

    path = int(eng.stat(_lib.STAT_PATH))
    alg = algorithmic_bytes(prob, nnz)     # this rank's launch (its sub-mesh incl. replicated halo elements)
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_file):
        peak, peak_src = json.load(open(peaks_file))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    # per-launch time of this rank's kernel: the timed region holds only numeric kernels
    achieved = alg / (ms_step / 1e3) / 1e9
    traffic = None
    tf = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tf):
        traffic = json.load(open(tf)).get(f"{args.workload}_N{n}")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": int(alg),
                "algorithmic_bytes_per_element": alg / prob.meshes[0].nel_,
                "designed_bytes_per_launch": int(eng.stat(_lib.STAT_NUMERIC_BYTES)),
                "frac_of_nominal_8TBps": achieved / 8000.0,
                "kernel": "k_tl_numeric" if path == 2 else "k_tp_elem_matrices+k_tp_gather"}

    # ---- SURVEY 8f rows f1 / f2, timed beside the hot path (heat workloads, one GPU): the load vector of the same
    # integrate! loop and K*T of the examples' solve!, device-resident, CUDA events inside the library ----------
    callers = None
    if world == 1 and prob.form.form_id == 1 and not args.no_callers and path == 2:
        m0 = prob.meshes[0]
        nd = int(prob.ndofs)
        eng.vec_assemble(_lib.VFORM_HEAT_LOAD, prob.quad, [-6.0], nd)          # builds the dof -> contribution map
        tv = []
        for _ in range(5):
            eng.vec_assemble(_lib.VFORM_HEAT_LOAD, prob.quad, [-6.0], nd)
            tv.append(eng.stat(_lib.STAT_VEC_MS))
        v_ms = float(np.median(tv))
        v_alg = 4 * m0.kind * m0.nel_ + 16 * m0.nnodes_ + 4 * nd + 4 * m0.kind * m0.nel_ + 8 * nd
        xd = torch.ones(nd, dtype=torch.float64, device="cuda")
        yd = torch.empty(nd, dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        eng.spmv(xd, yd)                                                       # builds the row-major view
        ts = []
        for _ in range(5):
            eng.spmv(xd, yd)
            ts.append(eng.stat(_lib.STAT_SPMV_MS))
        s_ms = float(np.median(ts))
        s_alg = 12 * nnz + 8 * nd + 8 * nd + 8 * (nd + 1)
        callers = {
            "load_vector": {"what": "efg_vec_assemble (SysvecAssembler: fe[j] += N[j]*Q*JxW), 2 kernels, dof map cached",
                            "ms": v_ms, "elements_per_s": m0.nel_ / (v_ms / 1e3), "algorithmic_bytes": int(v_alg),
                            "achieved_GBps": v_alg / (v_ms / 1e3) / 1e9, "frac": v_alg / (v_ms / 1e3) / 1e9 / peak},
            "spmv": {"what": "efg_spmv (KT = K*T, SparseArrays summation order); heat matrices are bitwise symmetric: CSC columns read in place as rows",
                     "ms": s_ms, "nnz_per_s": nnz / (s_ms / 1e3), "algorithmic_bytes": int(s_alg),
                     "achieved_GBps": s_alg / (s_ms / 1e3) / 1e9, "frac": s_alg / (s_ms / 1e3) / 1e9 / peak,
                     "y_checksum": float(yd.sum().item())},
        }
        del xd, yd

    out = {
        "metric": "elements assembled/s", "value": value, "unit": "elements/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{WORKLOADS[args.workload][1]}, N={n}" + (f", columns split over {world} ranks" if world > 1 else ""),
                   "elements": int(nel_global), "ndofs": int(prob.ndofs), "nnz": int(nnz_global), "l2": "inputs larger than L2",
                   "path": {1: "two-pass", 2: "tiled-fused"}[path], "strict_fp": args.strict,
                   "tile_elems": int(args.tile_elems) or -(-int(prob.meshes[0].nel_) // max(int(eng.stat(_lib.STAT_NTILES)), 1)),
                   "tile_elems_source": "option" if args.tile_elems else "automatic (largest size with two CTAs per SM)", "sfc_order": args.sfc,
                   "tiles": int(eng.stat(_lib.STAT_NTILES)),
                   "halo_factor": eng.stat(_lib.STAT_TILE_ELEMS) / max(prob.meshes[0].nel_, 1),
                   "rank_elements_incl_shard_halo": int(prob.meshes[0].nel_)},
        "roofline": roofline, "e2e": e2e, "gpu_launches": gpu_launches, "clocks": clocks,
        "phases": {"symbolic_ms": sym_ms, "numeric_ms": ms_step,
                   "value_with_symbolic": nel_global / ((ms_step + sym_ms) / 1e3)},
        "device_bytes": eng.stat(_lib.STAT_DEVICE_BYTES), "nzval_checksum": checksum, "next_rows": callers,
    }
    if rank == 0 and world == 1 and not args.no_cpu:
        cn = min(args.cpu_n, n)
        out["cpu_baseline"] = cpu_baseline(efg, args.workload, cn)
    eng.close()
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
