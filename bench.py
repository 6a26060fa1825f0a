#!/usr/bin/env python
"""bench.py - bootstrapped ResNet-20 CIFAR-10 homomorphic inference on the B200 engine (BASELINE.json's metric).

Workload (config.workload): the reference's `./cnn 20 10 i i` (ResNet_cifar10_seal_sparse, cnn_ckks/cpu-ckks/single-key/
cnn/infer_seal.cpp:251-584) at its own parameters - N = 2^16, primes 51 | 46x16 | 51x14 | 51, Hamming weight 192,
scale 2^46, 18 sparse-slot bootstraps, 19 multiplexed convolutions, 19 alpha=13 minimax ReLUs - on synthetic images
with the reference's trained parameters.  One step = `--in-flight` images per GPU.  Images are independent: image i goes
to rank i mod G, no data-path collective (weak scaling); rank 0 samples the secret key and broadcasts it with NCCL,
every rank derives its evaluation keys from it on its own GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl engine|reference] [--layers 20|110]

  value  images/s with the encrypted image already resident in HBM (device-timed, CUDA events, max over ranks)
  e2e    images/s through the C ABI call a user makes (bka_resnet_infer: host image in, host logits out; packing,
         encoding, encryption, every host->device copy of plaintext operands, decryption and decoding inside the
         timed region)
  roofline      the kernel family with the largest share of the step, timed live with CUDA events; int_roofline inside
                it is the same kernel against the integer multiply-add peak measured on this GPU in this run
  exact         the same workload with every tolerance-mode path switched off (key switching decomposed one digit per
                prime, rotations one by one, constants encrypted on the hot path: the reference's exact operation
                sequence, the mode the limb-level parity tests cover), measured in a child process of the same run
  key_switch_us rotate_vector at N = 2^16 and 31 / 17 / 3 limbs: this engine (both modes) and the reference per thread
  cpu_baseline  the reference's own SEAL (oracle/_ref/libseal_ref.so) on this host's cores: per-operation times
                measured per level on all cores, composed with the operation histogram of one inference
                (the full CPU run needs NTL, ~384 GB of RAM and ~2200 s per image: reference README / result log)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "fhe-gpt-2_b200", "python"))

LOG_N = 16
CNN_BITS = [51] + [46] * 16 + [51] * 14 + [51]   # infer_seal.cpp:288-311
LIMB_BYTES = (1 << LOG_N) * 8
METRIC = "ResNet-20 CIFAR-10 homomorphic inference throughput (bootstrapped, N=2^16)"
UNIT = "images/s"
HIST_KEYS = ["key_switch", "rescale", "multiply_vector", "multiply", "scalar", "add"]
NTT_BUTTERFLIES_PER_PASS = (1 << LOG_N) // 2 * 8      # a pass = 8 of the 16 stages of one limb-polynomial


def hist_path(layers):
    return os.path.join(ROOT, "tests", "golden", f"resnet{layers}_op_histogram.json")


def metric_name(layers):
    return METRIC if layers == 20 else METRIC.replace("ResNet-20", f"ResNet-{layers}")


def load_weights(layers):
    """the reference's trained parameters when the fixture is there (tests/golden/pretrained_parameters/resnet20_new,
    a copy of the reference's pretrained_parameters/resnet20_new), else random-init weights of the architecture"""
    from b200ckks import synthetic

    d = synthetic.pretrained_dir(layers)
    if d:
        return synthetic.load_pretrained(d, layers), f"trained parameters pretrained_parameters/resnet{layers}_new"
    return synthetic.random_weights(layers, seed=0), "random-init weights of the architecture"


def make_config(args):
    """identical for both arms (the driver compares the two config objects)"""
    from b200ckks import synthetic

    K = max(1, args.in_flight)
    weights = (f"trained parameters pretrained_parameters/resnet{args.layers}_new" if synthetic.pretrained_dir(args.layers)
               else "random-init weights of the architecture")
    return {"workload": f"ResNet-{args.layers} CIFAR-10 with bootstrapping (reference: ./cnn {args.layers} 10 i j); logN=16 RNS-CKKS, "
                        "primes 51|46x16|51x14|51, Hamming weight 192, scale 2^46",
            "log_n": LOG_N, "layers": args.layers, "images_per_step_per_unit": K,
            "images": "synthetic: N(0,1) clipped to [-2.5, 2.5], seed = image id (the reference's CIFAR image file is missing)",
            "weights": weights,
            "l2": "working set larger than L2: every bootstrap streams tens of GiB of evaluation keys and 31-limb "
                  "ciphertexts (31 MiB each) against a 126 MB L2",
            "parallelism": f"{args.gpus} unit(s), independent images dealt round-robin, no data-path exchange"}


# algorithmic bytes per limb-polynomial and kernel family: an NTT reads and writes a limb
# once (2 N w) and is executed as two passes, so each pass owns N w; the key-switch inner product streams one key limb
# per unit; element-wise kernels read and write their operand limbs.
ALGO_BYTES_PER_UNIT = {"fwd_cols": LIMB_BYTES, "fwd_blocks": LIMB_BYTES, "inv_blocks": LIMB_BYTES, "inv_cols": LIMB_BYTES,
                       "ks_mac": LIMB_BYTES, "elementwise": 2 * LIMB_BYTES, "fft": LIMB_BYTES, "other": LIMB_BYTES}


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
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
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
        sm, mx, reasons, power = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
                power.append(float(f[2]))
            except ValueError:
                continue
            for nme, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort()
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "samples": len(sm), "power_w_max": max(power) if power else None,
                "reasons": sorted(reasons)}


class stdout_to_stderr:
    """NCCL prints its version banner with printf on first use; stdout must carry the one JSON line only."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def dist_env():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")))


def collect_histogram(sess):
    return {k: sess.level_histogram(i) for i, k in enumerate(HIST_KEYS)}


def run_engine(args):
    import numpy as np
    import torch
    from b200ckks import synthetic
    from b200ckks.app import App

    rank, local, world = dist_env()
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        with stdout_to_stderr():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            warm = torch.zeros(1, device="cuda")
            dist.all_reduce(warm)                 # creates the communicator (and prints the banner) here
            torch.cuda.synchronize()

    # ---- keys: one secret for the node (rank 0 samples it, NCCL broadcast), evaluation keys derived per GPU ----
    os.environ.setdefault("B200CKKS_SEED", "0x5EA1C0DE")      # reproducible benchmark randomness (never set in production)
    hybrid = not args.no_hybrid
    os.environ["B200CKKS_HYBRID_KS"] = "1" if hybrid else "0"
    if args.compress_keys:
        os.environ["B200CKKS_COMPRESS_KEYS"] = "1"
    app = App()
    weights, weights_name = load_weights(args.layers)
    image_of = lambda step: synthetic.synthetic_image(rank + world * step)

    # Key plan (host/seal/seal.h KeyPlan): which (Galois element, level) keys the network touches.  It depends on the
    # network and the key-switching mode only, so it is read from fhe-gpt-2_b200/plans/ when committed there, else
    # learnt from one inference under a THROW-AWAY secret key.  The real session then generates exactly those keys up
    # front and detaches the secret key from every evaluation key: the timed region evaluates without it, as the
    # reference does after create_galois_keys (infer_seal.cpp:379).
    mode_name = "hybrid" if hybrid else "exact"
    plan_path = os.path.join(ROOT, "fhe-gpt-2_b200", "plans", f"resnet{args.layers}_{mode_name}.plan")
    t_plan = time.perf_counter()
    if args.lazy_keys:
        plan, plan_source = None, "none (--lazy-keys: keys generated on first use from the resident secret key)"
    elif os.path.exists(plan_path):
        plan, plan_source = open(plan_path).read(), os.path.relpath(plan_path, ROOT)
    else:
        dry = app.session(LOG_N, CNN_BITS, hamming_weight=192, device=local)
        dry_net = dry.resnet(args.layers, weights)
        dry_net.infer(image_of(0), trace=False)
        if args.in_flight > 1:      # the batch path takes the same rotations; run it once so nothing is missed
            dry_net.infer_batch(np.stack([image_of(0)] * 2), 2)
        plan, plan_source = dry.key_plan(), "dry run under a throw-away key"
        del dry_net
        dry.close()
        if rank == 0 and args.write_plan:
            os.makedirs(os.path.dirname(args.write_plan), exist_ok=True)
            open(args.write_plan, "w").write(plan)
    plan_s = time.perf_counter() - t_plan

    t_setup = time.perf_counter()
    sk_words = len(CNN_BITS) * (1 << LOG_N)
    if world > 1:
        buf = torch.empty(sk_words, dtype=torch.int64, device="cuda")
        if rank == 0:
            sess = app.session(LOG_N, CNN_BITS, hamming_weight=192, device=local)
            buf.copy_(torch.from_numpy(sess.secret_key().view(np.int64).reshape(-1)))
        dist.broadcast(buf, 0)
        torch.cuda.synchronize()
        if rank != 0:
            sk = buf.cpu().numpy().view(np.uint64).reshape(len(CNN_BITS), 1 << LOG_N)
            sess = app.session(LOG_N, CNN_BITS, hamming_weight=192, device=local, secret_key=sk)
        del buf
    else:
        sess = app.session(LOG_N, CNN_BITS, hamming_weight=192, device=local)
    eng = sess.engine()
    net = sess.resnet(args.layers, weights)
    t_gen = time.perf_counter()
    if plan is not None:
        sess.apply_key_plan(plan, detach_secret=True)
    gen_s = time.perf_counter() - t_gen
    keys_after_plan = sess.key_residency()[1]

    def barrier():
        sess.sync()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- warm-up: the first image also materialises the level-pruned Galois keys in HBM --------------------
    first_logits = None
    for w in range(args.warmup):
        logits, _ = net.infer(image_of(w), trace=False)
        if w == 0:
            first_logits = logits
    setup_s = time.perf_counter() - t_setup
    key_bytes, key_gens = sess.key_residency()
    if plan is not None and key_gens != keys_after_plan:
        raise SystemExit(f"rank {rank}: {key_gens - keys_after_plan} key(s) were generated during evaluation although the "
                         "keys come from a plan")
    barrier()

    # ---- value: encrypted images resident in HBM, device-timed -----------------------------------------------
    enc = [net.encrypt_image(image_of(args.warmup + k)) for k in range(args.steps)]
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    sess.stats(reset=True)
    for i in range(len(HIST_KEYS)):
        sess.level_histogram(i, reset=True)
    launches0 = eng.launch_count()
    outs = []
    torch.cuda.profiler.start()      # cudaProfilerStart: `ncu --profile-from-start off` captures the timed region only
    eng.timer_begin()
    for k in range(args.steps):
        outs.append(net.infer_encrypted(enc[k])[0])
    ms = eng.timer_end()
    torch.cuda.profiler.stop()
    launches = eng.launch_count() - launches0
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    stats = sess.stats()
    hist = collect_histogram(sess)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    logits_timed = [net.decrypt_logits(o) for o in outs]
    if not all(np.isfinite(l).all() and np.abs(l).max() < 1e3 for l in logits_timed):
        raise SystemExit(f"rank {rank}: the encrypted network produced non-finite or exploded logits")
    latency_ms = ms
    del outs

    # ---- throughput: K images in flight per GPU (one host thread + CUDA stream each, shared keys), the reference's
    # own parallel unit (`#pragma omp parallel for` over images, infer_seal.cpp:404).  A step = K images per GPU.
    K = max(1, args.in_flight)
    launches_tp = launches
    if K > 1:
        enc_tp = [enc[k % args.steps].clone() for k in range(args.steps * K)]
        net.infer_encrypted_batch(enc_tp[:K], K)         # worker threads, their streams and staging rings (untimed)
        barrier()
        launches0 = eng.launch_count()
        eng.timer_begin()
        outs_tp = net.infer_encrypted_batch(enc_tp, K)
        ms = eng.timer_end()
        launches_tp = eng.launch_count() - launches0
        barrier()
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if dist:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        tp_logits = [net.decrypt_logits(o) for o in outs_tp]
        drift_tp = max(float(np.abs(l - logits_timed[k % args.steps]).max()) for k, l in enumerate(tp_logits))
        if not drift_tp < 5e-2:
            raise SystemExit(f"rank {rank}: concurrent and sequential inference disagree (max |dlogit| {drift_tp})")
        del enc_tp, outs_tp
    del enc

    # ---- e2e: the user-facing call with host buffers; every copy inside the timed region ---------------------
    e2e_steps = args.steps
    n_e2e = e2e_steps * K
    pinned = torch.empty((n_e2e, 3072), dtype=torch.float64).pin_memory()
    for k in range(n_e2e):
        pinned[k].copy_(torch.from_numpy(image_of(args.warmup + k % args.steps)))
    barrier()
    h2d0, d2h0 = eng.transfer_bytes()
    t0 = time.perf_counter()
    if K > 1:
        e2e_logits = list(net.infer_batch(pinned.numpy(), K))
    else:
        e2e_logits = [net.infer(pinned[k].numpy(), trace=False)[0] for k in range(n_e2e)]
    sess.sync()
    e2e_s = time.perf_counter() - t0
    h2d1, d2h1 = eng.transfer_bytes()
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    # the same images went through both paths: their logits must agree up to encryption noise
    drift = max(float(np.abs(a - logits_timed[k % args.steps]).max()) for k, a in enumerate(e2e_logits))
    if not drift < 5e-2:
        raise SystemExit(f"rank {rank}: the two timed paths disagree on the same images (max |dlogit| {drift})")

    # ---- roofline of the dominant kernel family: one more image with CUDA events around every launch ----------
    roof, kernels = None, None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        which = "measured (MEASURED_PEAKS.json hbm_gbs, sustained copy)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
        ct = net.encrypt_image(image_of(0))
        sess.sync()
        c0 = eng.kernel_counters()
        eng.profile_begin_all()
        eng.timer_begin()
        net.infer_encrypted(ct)
        prof_ms = eng.timer_end()
        prof = eng.profile_end_all()
        c1 = eng.kernel_counters()
        total_kernel_ms = sum(v[1] for v in prof.values())
        kernels = {}
        for fam, (n_launch, fam_ms) in prof.items():
            units = c1[fam][1] - c0[fam][1]
            if n_launch == 0:
                continue
            gbs = units * ALGO_BYTES_PER_UNIT[fam] / (fam_ms * 1e-3) / 1e9 if fam_ms > 0 else 0.0
            kernels[fam] = {"launches": n_launch, "ms": round(fam_ms, 2), "share_of_kernel_time": round(fam_ms / total_kernel_ms, 4),
                            "limb_polys": units, "algorithmic_GBps": round(gbs, 1), "frac_of_hbm_peak": round(gbs / peak, 4)}
        dom = max(kernels, key=lambda k: kernels[k]["ms"])
        # DRAM bytes of this kernel from an `ncu --set full` capture.  ncu over the ~94k launches of an image is
        # impractical, so the capture is the same kernel inside an l = 31 key switch (its dominant caller here); the
        # algorithmic bytes of THAT launch are given beside it (traffic / traffic_algorithmic_bytes = re-read factor).
        traffic, traffic_algo, traffic_src = None, None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = tj.get(f"{dom}_resnet20", tj.get(f"{dom}_l31"))
            traffic_algo = tj.get(f"{dom}_resnet20_algorithmic", tj.get(f"{dom}_l31_algorithmic"))
            traffic_src = tj.get(f"{dom}_resnet20_source", "profiles/r1_ncu_full_keyswitch_l31.md (one launch of this kernel in an "
                                                           "l=31 key-switch chunk)")
        except Exception:
            pass
        d = kernels[dom]
        roof = {"bound": "hbm", "kernel": {"fwd_cols": "k_fwd_cols (NTT column pass)", "fwd_blocks": "k_fwd_blocks (NTT block pass)",
                                            "inv_blocks": "k_inv_blocks", "inv_cols": "k_inv_cols", "ks_mac": "k_ks_mac",
                                            "elementwise": "k_ew / k_scalar_pack", "fft": "k_fft_*",
                                            "other": "k_hyb_conv (basis conversion) and samplers"}[dom],
                "achieved": d["algorithmic_GBps"], "peak": peak, "unit": "GB/s", "frac": d["frac_of_hbm_peak"], "traffic": traffic,
                "traffic_algorithmic_bytes": traffic_algo,
                "traffic_source": traffic_src,
                "peak_source": which, "launches_per_step": d["launches"], "ms_per_launch": round(d["ms"] / d["launches"], 5),
                "kernel_share_of_step": round(d["ms"] / prof_ms, 4),
                "algorithmic_bytes_per_launch": int(d["limb_polys"] * ALGO_BYTES_PER_UNIT[dom] / d["launches"]),
                "note": "the NTT passes are bound by the integer multiply-add pipe, not by HBM (ncu captures in profiles/); the "
                        "HBM fraction is reported because the contract asks for hbm|tensor, int_roofline is the bound that applies",
                "gpu_busy_fraction_of_step": round(total_kernel_ms / prof_ms, 4)}
        # Integer roofline, measured in this run.  The multiplier pipe of an SM issues the three forms a 64-bit Shoup
        # butterfly is made of at different rates (bk_measure_int_pipe: dependent chains of mad.lo / mad.wide / mad.hi
        # on this GPU); the unreduced forward butterfly (ntt.cuh ct_bfly_wide) compiles to 4 IMAD + 3 IMAD.WIDE +
        # 2 IMAD.HI, the Gentleman-Sande butterfly of the inverse (exact 64-bit high product) to 4 + 6 + 0 (executed
        # instruction mix of the ncu captures in profiles/r2_ntt_lab.md).  frac = multiplier-pipe seconds those instructions need at the measured rates /
        # live-timed seconds of the family (2^15 x 8 butterflies per pass and limb-polynomial, limb-polynomials from
        # the engine's kernel counters).  1.0 would be a pass that issues nothing but its butterflies' multiplies.
        rates = eng.int_pipe_rates()
        mix = {"fwd": {"mad_lo": 4, "mad_wide": 3, "mad_hi": 2}, "inv": {"mad_lo": 4, "mad_wide": 6, "mad_hi": 0}}
        pipe_s = {k: sum(n / rates[f] for f, n in m.items()) for k, m in mix.items()}   # thread-seconds per butterfly
        int_roof = {"rates_per_s": {k: round(v) for k, v in rates.items()},
                    "rates_source": "bk_measure_int_pipe: 8 independent dependent chains per thread of each form, this GPU, this run",
                    "instructions_per_butterfly": mix, "butterflies_per_limb_poly_pass": NTT_BUTTERFLIES_PER_PASS,
                    "peak_butterflies_per_s": {k: round(1.0 / v) for k, v in pipe_s.items()}}
        for fam in ("fwd_cols", "fwd_blocks", "inv_blocks", "inv_cols"):
            if fam in kernels and kernels[fam]["ms"] > 0:
                bf = kernels[fam]["limb_polys"] * NTT_BUTTERFLIES_PER_PASS / (kernels[fam]["ms"] * 1e-3)
                int_roof[fam] = {"achieved_butterflies_per_s": round(bf), "frac": round(bf * pipe_s[fam[:3]], 4)}
        if dom in int_roof:
            int_roof["frac"] = int_roof[dom]["frac"]
        roof["int_roofline"] = int_roof

    # ---- BASELINE config 4 beside it (rank 0, N = 1, default workload only): ResNet-110, 109 convolutions and 108
    # bootstraps, same session and keys (the key plan covers it: same layer shapes, more of them), one image alone,
    # random-init weights of the architecture (the fixture holds resnet20_new only) ------------------------------
    resnet110 = None
    if rank == 0 and world == 1 and args.layers == 20 and not args.no_resnet110:
        try:
            from b200ckks import synthetic

            net110 = sess.resnet(110, synthetic.random_weights(110, seed=0))
            img110 = image_of(0)
            net110.infer(img110, trace=False)            # untimed: encodes and caches its 109 layers' plaintexts
            ct110 = [net110.encrypt_image(img110) for _ in range(2)]
            sess.sync()
            eng.timer_begin()
            out110 = [net110.infer_encrypted(c)[0] for c in ct110]
            ms110 = eng.timer_end()
            l110 = net110.decrypt_logits(out110[0])
            resnet110 = {"seconds_per_image": round(ms110 * 1e-3 / len(ct110), 3), "images_in_flight": 1, "bootstraps": 108,
                         "weights": "random-init weights of the architecture", "logits_finite": bool(np.isfinite(l110).all()),
                         "reference": "./cnn 110 10 i i (BASELINE.json configs[3])"}
            del net110, out110, ct110
        except Exception as e:  # a key outside the plan, or memory: report it rather than lose the line
            resnet110 = {"unavailable": str(e)[:200]}

    # ---- CPU baseline beside it (rank 0, N = 1 only; bounded sample) ---------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        per_image = {k: [c / args.steps for c in v] for k, v in hist.items()}
        cpu = cpu_baseline(per_image, reps=1)

    # ---- key-switch microseconds and the exact mode (rank 0, N = 1 only): the main session is closed first ------
    ks_us, exact = None, None
    plain_cache = sess.plain_cache()
    if rank == 0 and world == 1 and not args.no_exact:
        del net
        sess.close()
        ks_us = {"engine_hybrid" if hybrid else "engine_exact": key_switch_us(app, local)}
        if cpu and cpu.get("key_switch_seconds_per_thread"):
            ks_us["reference_per_thread"] = {str(l): round(v * 1e6, 1) for l, v in cpu["key_switch_seconds_per_thread"].items()
                                             if int(l) in (31, 17, 3)}
        if hybrid:
            exact = run_exact_child(args)
            if exact and "key_switch_us" in exact:
                ks_us["engine_exact"] = exact.pop("key_switch_us")

    if rank == 0:
        total_images = world * args.steps * K
        line = {
            "metric": metric_name(args.layers), "value": total_images / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": f"synthetic images; {weights_name}",
            "config": make_config(args),
            "engine": {"images_in_flight_per_gpu": K,
                       "mode": ("tolerance mode: level-aware hybrid key switching, hoisted baby-step rotations, constants as "
                                "plaintexts (decrypted values equal the reference's, limbs do not)" if hybrid
                                else "key switching decomposed one digit per prime as in the reference"),
                       "parallelism": f"dp{world}: image i -> rank i mod {world}; secret key broadcast once over NCCL, evaluation keys "
                                      f"generated per GPU; no data-path collective; per GPU {K} host threads / CUDA streams, one "
                                      "image each, over shared keys (the reference's OpenMP image loop, infer_seal.cpp:404)"},
            "seconds_per_image": ms * 1e-3 / (args.steps * K),
            "latency": {"seconds_per_image": latency_ms * 1e-3 / args.steps, "images_in_flight": 1,
                        "note": "one image alone on the GPU, device-timed over `steps` images"},
            "e2e": {"value": world * n_e2e / e2e_s, "unit": UNIT, "seconds_per_image": e2e_s / n_e2e,
                    "h2d_bytes_per_step": (h2d1 - h2d0) // e2e_steps, "d2h_bytes_per_step": (d2h1 - d2h0) // e2e_steps,
                    "call": ("bka_resnet_infer_batch(net, images[n][3072] on the host, in_flight) -> logits[n][10] on the host" if K > 1
                             else "bka_resnet_infer(net, image[3072] on the host) -> logits[10] on the host")},
            "exact": exact, "key_switch_us": ks_us, "resnet110": resnet110,
            "gpu_launches": int(launches_tp), "clocks": clocks, "roofline": roof, "kernels": kernels, "cpu_baseline": cpu,
            "ops_per_image": {k: v // args.steps for k, v in stats.items()},
            "keys": {"plan": plan_source, "plan_seconds": round(plan_s, 1), "generate_seconds": round(gen_s, 1),
                     "secret_key_detached_from_evaluation_keys": plan is not None,
                     "resident_gib": round(key_bytes / 2 ** 30, 2), "seed_compressed": bool(args.compress_keys), "generated": key_gens,
                     "generated_during_evaluation": (key_gens - keys_after_plan) if plan is not None else key_gens,
                     "setup_and_warmup_seconds": round(setup_s, 1)},
            "plaintext_cache": {k: (round(v / 2 ** 30, 2) if k == "bytes" else v) for k, v in plain_cache.items()},
            "check": {"logits_image0": [round(float(x), 4) for x in first_logits],
                      "max_logit_difference_between_timed_paths": drift},
        }
        print(json.dumps(line), flush=True)
        if args.dump_histogram:
            json.dump({k: [int(round(c / args.steps)) for c in v] for k, v in hist.items()}, open(args.dump_histogram, "w"))
    if dist:
        dist.barrier()
        dist.destroy_process_group()


def key_switch_us(app, device, reps=10):
    """rotate_vector(step 1) at N = 2^16 and 31 / 17 / 3 limbs, microseconds (CUDA events, mean of `reps` after 3 warm-ups),
    in the key-switching mode of the current environment"""
    import numpy as np

    s = app.session(LOG_N, CNN_BITS, hamming_weight=192, device=device, rotation_steps=[1])
    eng = s.engine()
    out = {}
    x = np.linspace(-1, 1, s.slots)
    for limbs in (31, 17, 3):
        ct = s.encrypt(x, 2.0 ** 46, limbs=limbs)
        for _ in range(3):
            s.rotate(ct, 1)
        s.sync()
        eng.timer_begin()
        for _ in range(reps):
            s.rotate(ct, 1)
        out[str(limbs)] = round(eng.timer_end() * 1e3 / reps, 1)
    s.close()
    return out


def run_exact_child(args):
    """the exact mode in a process of its own (its switches are read once per process)"""
    env = dict(os.environ, B200CKKS_HYBRID_KS="0", B200CKKS_NO_HOIST="1", B200CKKS_ENCRYPT_CONSTANTS="1")
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "exact-child", "--layers", str(args.layers), "--steps", str(min(args.steps, 3)),
           "--in-flight", str(args.in_flight)]
    try:
        r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=1500)
        lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if r.returncode != 0 or not lines:
            return {"error": (r.stderr or r.stdout)[-400:]}
        return json.loads(lines[-1])
    except Exception as e:      # the headline must not die with the side measurement
        return {"error": repr(e)}


def run_exact_child_body(args):
    """ResNet with every tolerance-mode path off: one digit per prime, rotations one by one, constants encrypted on the
    hot path - the reference's exact operation sequence (the limb-level parity tests cover this mode)"""
    import numpy as np
    from b200ckks import synthetic
    from b200ckks.app import App

    app = App()
    os.environ.setdefault("B200CKKS_SEED", "0x5EA1C0DE")
    weights, _ = load_weights(args.layers)
    sess = app.session(LOG_N, CNN_BITS, hamming_weight=192)
    eng = sess.engine()
    net = sess.resnet(args.layers, weights)
    img = lambda k: synthetic.synthetic_image(k)
    first = net.infer(img(0), trace=False)[0]          # warm-up: materialises the level-pruned keys
    enc = [net.encrypt_image(img(1 + k)) for k in range(args.steps)]
    sess.sync()
    eng.timer_begin()
    outs = [net.infer_encrypted(c)[0] for c in enc]
    alone_ms = eng.timer_end()
    K = max(1, args.in_flight)
    batch = [enc[k % args.steps].clone() for k in range(K * 2)]
    net.infer_encrypted_batch(batch[:K], K)
    sess.sync()
    eng.timer_begin()
    net.infer_encrypted_batch(batch, K)
    tp_ms = eng.timer_end()
    key_bytes, key_gens = sess.key_residency()
    logits = net.decrypt_logits(outs[0])
    if not (np.isfinite(logits).all() and np.abs(logits).max() < 1e3):
        raise SystemExit("exact mode produced non-finite logits")
    del net, outs, enc, batch
    sess.close()
    res = {"mode": "reference operation sequence: key switching one digit per prime, rotations one by one, constants encrypted "
                   "(B200CKKS_HYBRID_KS=0 B200CKKS_NO_HOIST=1 B200CKKS_ENCRYPT_CONSTANTS=1)",
           "value": (K * 2) / (tp_ms * 1e-3), "unit": UNIT, "images_in_flight": K,
           "seconds_per_image": tp_ms * 1e-3 / (K * 2), "latency_seconds_per_image": alone_ms * 1e-3 / args.steps,
           "galois_keys_resident_gib": round(key_bytes / 2 ** 30, 2),
           "logits_image0": [round(float(x), 4) for x in first],
           "key_switch_us": key_switch_us(app, 0)}
    print(json.dumps(res), flush=True)


# ---------------------------------------------------------------------------------------------------- CPU arm
def host_threads():
    """cores this process may run on (torchrun exports OMP_NUM_THREADS=1, which the explicit num_threads clause of the
    timing loop overrides; the affinity mask is what really bounds the OpenMP team)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def measure_reference_ops(threads, reps):
    """Per-operation seconds of the reference's own SEAL at the CNN parameters, `threads` independent ciphertexts at a
    time (how infer_seal.cpp:404 parallelises), at a few levels; returns {op: {limbs: seconds}}."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import refseal

    ref = refseal.RefSeal(LOG_N, CNN_BITS, hamming_weight=192, seed=7)
    ref.make_galois_keys([1])
    ref.relin_key(dump=False)
    pt, top = ref.pt_new(), ref.ct_new()
    ref.encode(pt, np.linspace(-1, 1, 1 << (LOG_N - 1)), 31, 2.0 ** 46)
    ref.encrypt(pt, top)
    out = {k: {} for k in HIST_KEYS}
    for l in (31, 24, 17, 10, 3):
        a = ref.ct_new()
        ref.ct_copy(a, top)
        if l < 31:
            ref.op("mod_switch_to", a, iarg=l)
        out["key_switch"][l] = ref.time_op("rotate", a, iarg=1, threads=threads, reps=reps)[1]
        if l in (31, 17, 3):
            out["rescale"][l] = ref.time_op("rescale", a, threads=threads, reps=reps)[1]
            out["multiply_vector"][l] = ref.time_op("multiply_vector", a, darg=0.5, threads=threads, reps=reps)[1]
            out["multiply"][l] = ref.time_op("multiply", a, b=a, threads=threads, reps=reps)[1]
            out["scalar"][l] = ref.time_op("multiply_const", a, darg=0.5, threads=threads, reps=reps)[1]
            out["add"][l] = ref.time_op("add", a, b=a, threads=threads, reps=reps)[1]
        ref.ct_free(a)
    ref.close()
    return out


def compose_seconds_per_image(per_op, hist):
    """sum over operation classes and levels of count x measured seconds (piecewise-linear in the limb count)."""
    import numpy as np

    total, parts = 0.0, {}
    for op, counts in hist.items():
        xs = sorted(per_op[op])
        ys = [per_op[op][x] for x in xs]
        s = 0.0
        for limbs, c in enumerate(counts):
            if c:
                s += c * float(np.interp(limbs, xs, ys))
        parts[op] = s
        total += s
    return total, parts


def cpu_baseline(hist, reps=1, threads=None):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import refseal

    if not refseal.available():
        return {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "oracle/_ref not built"}
    T = threads or host_threads()
    per_op = measure_reference_ops(T, reps)
    sec, parts = compose_seconds_per_image(per_op, hist)
    return {"value": T / sec, "unit": UNIT, "cores": T, "kind": "reference", "seconds_per_image_per_thread": sec,
            "seconds_by_operation": {k: round(v, 1) for k, v in parts.items()},
            "key_switch_seconds_per_thread": {str(l): v for l, v in per_op["key_switch"].items()},
            "sample": f"the reference's modified SEAL 3.6.6 (compiled in place, -O2, no HEXL): rotate / rescale / multiply_vector / "
                      f"multiply / multiply_const / add timed at 3-5 levels with {T} OpenMP threads each on its own ciphertext "
                      f"({reps} call(s) per thread and level), composed with the per-level operation counts of one "
                      f"inference of this run; the reference's own single-thread log reports 2188.8 s per image for ResNet-20 "
                      f"(result/resnet20_cifar10_image0.txt) and its full run needs ~384 GB of RAM"}


def run_reference(args):
    rank, _, _ = dist_env()
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import refseal

    if not refseal.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libseal_ref.so is not built"}))
        return
    if not os.path.exists(hist_path(args.layers)):
        print(json.dumps({"impl": "reference", "unavailable": "no committed operation histogram for this depth"}))
        return
    hist = json.load(open(hist_path(args.layers)))
    T = host_threads()
    for _ in range(args.warmup):
        measure_reference_ops(T, 1)
    t0 = time.perf_counter()
    secs = []
    for _ in range(args.steps):
        secs.append(compose_seconds_per_image(measure_reference_ops(T, 1), hist)[0])
    wall = time.perf_counter() - t0
    sec = sum(secs) / len(secs)
    value = T / sec
    sample = (f"each step measures the reference's modified SEAL 3.6.6 operations (rotate, rescale, multiply_vector, multiply, "
              f"multiply_const, add) at 3-5 levels of the CNN chain with {T} OpenMP threads, one ciphertext per thread, and "
              f"composes them with the committed per-level operation counts of one ResNet-20 inference "
              f"(tests/golden/resnet{args.layers}_op_histogram.json); value = {T} concurrent images / composed seconds per image")
    weights_name = make_config(args)["weights"]
    line = {"impl": "reference", "metric": metric_name(args.layers), "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": wall * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": f"synthetic images; {weights_name}",
            "config": make_config(args), "threads": T,
            "seconds_per_image": sec / T, "seconds_per_image_per_thread": sec,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": T, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference", "exact-child"])
    ap.add_argument("--layers", type=int, default=20)
    ap.add_argument("--no-hybrid", action="store_true",
                    help="key switching exactly as the reference decomposes it (one digit per prime); default: level-aware "
                         "hybrid key switching (tolerance mode)")
    ap.add_argument("--in-flight", type=int, default=4, help="images in flight per GPU (a step = that many images per GPU)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-resnet110", action="store_true", help="skip the ResNet-110 leg (BASELINE config 4, one image alone)")
    ap.add_argument("--no-exact", action="store_true", help="skip the exact-mode child run and the key-switch timings")
    ap.add_argument("--compress-keys", action="store_true",
                    help="seed-compressed evaluation keys: the uniform half of every level key is regenerated from its public "
                         "seed when the key is used instead of being resident (half the key bytes in HBM)")
    ap.add_argument("--lazy-keys", action="store_true",
                    help="round-1 behaviour: evaluation keys generated on first use from the resident secret key instead of "
                         "up front from a key plan")
    ap.add_argument("--write-plan", default=None, help="write the key plan learnt from the dry run to this file")
    ap.add_argument("--dump-histogram", default=None, help="write the per-level operation counts of one inference (JSON)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "exact-child":
        run_exact_child_body(args)
    else:
        args.warmup = max(args.warmup, 3)
        run_engine(args)


if __name__ == "__main__":
    main()
