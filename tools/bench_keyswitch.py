#!/usr/bin/env python
"""bench.py - hot-path benchmark of the B200-native RNS-CKKS engine (contract: see DESIGN.md §5).

Workload (config.workload): `rotate_vector` = Galois automorphism + SEAL key switch
(evaluator.cpp:2224-2279 -> :2120-2222 -> :2281-2525 in the reference) on a batch of B
ciphertexts at N = 2^16, l = 31 limbs, the CNN prime chain {51, 46x16, 51x14 | 51} - the
"key-switch us at N=2^16" half of BASELINE.json's metric, and the operation that is >70 % of a
bootstrapped ResNet-20 inference.  One step = one pass of the key switch over the batch.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl engine|reference] [--limbs L] [--batch B]

N > 1 is launched with torch.distributed.run, one rank per GPU; the evaluation key is generated
once on rank 0 and broadcast over NCCL (keys generated once, BASELINE north_star); ciphertext
batches are independent per rank (weak scaling, no data-path collective).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fhe-gpt-2_b200", "python"))

LOG_N = 16
CNN_BITS = [51] + [46] * 16 + [51] * 14 + [51]   # infer_seal.cpp:288-322 of the reference
SCALE = 2.0 ** 46
LIMB_BYTES = (1 << LOG_N) * 8
METRIC = "key-switch throughput (rotate_vector) at N=2^16"
UNIT = "keyswitch/s"


def algorithmic_bytes_keyswitch(l):
    """target l + key 2l(l+1) + ciphertext read-modify-write 4l limb-polys."""
    return (2 * l * l + 7 * l) * LIMB_BYTES


def algorithmic_bytes_mac(l):
    """dominant kernel (k_ks_mac): the key stream 2l(l+1) limb-polys + accumulators 2(l+1)."""
    return (2 * l * (l + 1) + 2 * (l + 1)) * LIMB_BYTES


class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nme, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort()
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "samples": len(sm), "reasons": sorted(reasons)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    return rank, local, world


def run_engine(args):
    import numpy as np
    import torch
    import b200ckks as bk

    rank, local, world = dist_env()
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    l, B = args.limbs, args.batch
    primes = bk.coeff_modulus_create(LOG_N, CNN_BITS)
    eng = bk.Context(LOG_N, primes, device=local)
    elt = bk.galois_elt_from_step(LOG_N, 1)

    # ---- keys: generated once (rank 0), broadcast over NCCL ------------------------------------
    key_words = 31 * 2 * 32 * (1 << LOG_N)
    sk_buf = torch.empty(32 * (1 << LOG_N), dtype=torch.int64, device="cuda")
    key_buf = torch.empty(key_words, dtype=torch.int64, device="cuda") if world > 1 else None
    if rank == 0:
        sk = eng.generate_secret_key(192, seed=1)
        gkey = eng.create_galois_key(sk, elt, seed=4)
        if world > 1:
            gkey.export_device(key_buf.data_ptr())
            sk_buf.copy_(torch.from_numpy(sk.download().view(np.int64).reshape(-1)))
    if world > 1:
        torch.cuda.synchronize()
        dist.broadcast(key_buf, 0)
        dist.broadcast(sk_buf, 0)
        torch.cuda.synchronize()
        if rank != 0:
            gkey = eng.import_kskey_device(key_buf.data_ptr(), 31, 31)
            sk = eng.upload_secret_key(sk_buf.cpu().numpy().view(np.uint64).reshape(32, -1))
        del key_buf
    gk = eng.galois_keys()
    gk.set(elt, gkey)
    pk = eng.create_public_key(sk, seed=2 + rank)

    # ---- inputs resident in HBM ---------------------------------------------------------------------
    rng = np.random.default_rng(1234 + rank)
    x = rng.uniform(-1, 1, (B, 1 << (LOG_N - 1)))
    cts = []
    for b in range(B):
        ct = eng.encrypt(pk, eng.encode(x[b], 31, SCALE), seed=100 + b)
        eng.mod_switch_to_inplace(ct, l)
        cts.append(ct)
    eng.sync()

    def step():
        for ct in cts:
            eng.rotate_vector_inplace(ct, 1, gk)

    def barrier():
        eng.sync()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count()
    eng.timer_begin()
    for _ in range(args.steps):
        step()
    ms = eng.timer_end()
    launches = eng.launch_count() - launches0
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())

    # correctness of what was just timed: total rotation of ct 0 = warmup + steps slots
    total_rot = args.warmup + args.steps
    got = eng.decode(eng.decrypt(sk, cts[0]))
    err = float(np.max(np.abs(got - np.roll(x[0], -total_rot))))
    if not err < 1e-4:
        raise SystemExit(f"rank {rank}: rotated ciphertext decrypts wrong (max err {err})")

    # ---- e2e: host buffers through the C ABI, copies inside the timed region ---------------
    words = 2 * l * (1 << LOG_N)
    host_in = torch.empty((B, words), dtype=torch.int64).pin_memory()
    host_out = torch.empty((B, words), dtype=torch.int64).pin_memory()
    for b in range(B):
        cts[b].download_ptr(host_in[b].data_ptr())
    work = [eng.ciphertext() for _ in range(B)]

    def e2e_step():
        for b in range(B):
            work[b].upload_ptr(host_in[b].data_ptr(), 2, l, SCALE, True)
            eng.rotate_vector_inplace(work[b], 1, gk)
            work[b].download_ptr(host_out[b].data_ptr())

    e2e_steps = max(1, min(args.steps, 5))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    eng.sync()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())

    # ---- roofline of the dominant kernel, measured live with CUDA events ---------------------
    roof = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        which = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
        eng.flush_l2()
        eng.profile_begin("ks_mac")
        step()
        n_launch, mac_ms = eng.profile_end()
        per_ks_mac_ms = mac_ms / B
        achieved = algorithmic_bytes_mac(l) / (per_ks_mac_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get(f"ks_mac_l{l}")
            except Exception:
                traffic = None
        roof = {"bound": "hbm", "kernel": "k_ks_mac", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": which,
                "launches_per_keyswitch": n_launch / B, "ms_per_launch": mac_ms / max(n_launch, 1),
                "kernel_share_of_step": round(mac_ms / (ms / args.steps), 4),
                "algorithmic_bytes_per_keyswitch_kernel": algorithmic_bytes_mac(l),
                "op": {"algorithmic_bytes": algorithmic_bytes_keyswitch(l),
                       "achieved": round(algorithmic_bytes_keyswitch(l) / (ms * 1e-3 / (args.steps * B)) / 1e9, 1),
                       "frac": round(algorithmic_bytes_keyswitch(l) / (ms * 1e-3 / (args.steps * B)) / 1e9 / peak, 4)}}

    # ---- CPU baseline beside it (rank 0, N = 1 only; bounded sample) ---------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(l, reps=1)

    if rank == 0:
        total_ks = world * B * args.steps
        line = {
            "metric": METRIC, "value": total_ks / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": f"rotate_vector(step=1) = Galois permutation + key switch, batch of {B} "
                                   f"ciphertexts per GPU, N=2^16, l={l} limbs, CNN chain 51|46x16|51x14|51 "
                                   f"(ResNet-20 parameter set); Hamming-weight-192 secret",
                       "batch_per_gpu": B, "limbs": l, "log_n": LOG_N,
                       "l2": "inputs larger than L2: each key switch streams a 992 MiB evaluation key and "
                             f"the batch is {B * 2 * l * LIMB_BYTES >> 20} MiB",
                       "parallelism": f"dp{world} (independent ciphertexts per GPU, key broadcast once over NCCL)"},
            "us_per_keyswitch": ms * 1e3 / (B * args.steps),
            "e2e": {"value": world * B * e2e_steps / e2e_s, "unit": UNIT,
                    "h2d_bytes_per_step": B * words * 8, "d2h_bytes_per_step": B * words * 8,
                    "us_per_keyswitch": e2e_s * 1e6 / (B * e2e_steps)},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
            "check": {"decrypt_max_err_after_rotations": err},
        }
        print(json.dumps(line), flush=True)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(l, reps=1, threads=None):
    """The reference's own SEAL (oracle/_ref/libseal_ref.so) timed on this host's cores, one
    ciphertext per OpenMP thread (infer_seal.cpp:404 style)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import refseal

    if not refseal.available():
        return {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "oracle/_ref not built"}
    ref = refseal.RefSeal(LOG_N, CNN_BITS, hamming_weight=192, seed=7)
    ref.make_galois_keys([1])
    pt, ct = ref.pt_new(), ref.ct_new()
    ref.encode(pt, np.linspace(-1, 1, 1 << (LOG_N - 1)), 31, SCALE)
    ref.encrypt(pt, ct)
    if l < 31:
        ref.op("mod_switch_to", ct, iarg=l)
    T = threads or min(ref.max_threads(), os.cpu_count() or 1)
    wall, mean = ref.time_op("rotate", ct, iarg=1, threads=T, reps=reps)
    out = {"value": T * reps / wall, "unit": UNIT, "cores": T, "kind": "reference",
           "single_thread_us_per_keyswitch": mean * 1e6,
           "sample": f"{T * reps} rotate_vector calls at l={l} ({reps} per thread on {T} OpenMP threads), "
                     f"the reference's modified SEAL 3.6.6 compiled -O2 without HEXL"}
    ref.close()
    return out


def run_reference(args):
    rank, local, world = dist_env()
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import refseal

    l = args.limbs
    if not refseal.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libseal_ref.so is not built"}))
        return
    ref = refseal.RefSeal(LOG_N, CNN_BITS, hamming_weight=192, seed=7)
    ref.make_galois_keys([1])
    pt, ct = ref.pt_new(), ref.ct_new()
    ref.encode(pt, np.linspace(-1, 1, 1 << (LOG_N - 1)), 31, SCALE)
    ref.encrypt(pt, ct)
    if l < 31:
        ref.op("mod_switch_to", ct, iarg=l)
    T = min(ref.max_threads(), os.cpu_count() or 1)
    for _ in range(args.warmup):
        ref.time_op("rotate", ct, iarg=1, threads=T, reps=1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref.time_op("rotate", ct, iarg=1, threads=T, reps=1)
    dt = time.perf_counter() - t0
    value = T * args.steps / dt
    sample = (f"each step = {T} rotate_vector calls (one per OpenMP thread, independent ciphertexts) at l={l}; "
              f"the reference's modified SEAL 3.6.6 compiled in place -O2, no HEXL")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3 / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": f"rotate_vector(step=1), N=2^16, l={l} limbs, CNN chain 51|46x16|51x14|51",
                       "limbs": l, "log_n": LOG_N, "batch_per_step": T},
            "us_per_keyswitch": dt * 1e6 / (T * args.steps),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": T, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--limbs", type=int, default=31)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "engine" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_engine(args)


if __name__ == "__main__":
    main()
