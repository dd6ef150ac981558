#!/usr/bin/env python3
"""Digest of the scaling runs of a round: reads gpurun_out/<tag>{bench,corpus_dev,corpus_host,sweep_rows}_<N>gpu.json
(written by tools/run_multi.sh N <tag>), copies them to profiles/<out>* and writes profiles/<out>scaling.json.

    python tools/scaling_digest.py r2f_ r2_
"""
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, out = sys.argv[1], sys.argv[2]
src = os.path.join(ROOT, "gpurun_out")
dst = os.path.join(ROOT, "profiles")


def last_json(path):
    with open(path) as f:
        lines = [l for l in f.read().strip().splitlines() if l.startswith("{") or l.startswith("[")]
    return json.loads(lines[-1])


rows = []
for n in (1, 2, 4, 8):
    names = {k: f"{tag}{k}_{n}gpu.json" for k in ("bench", "corpus_dev", "corpus_host", "sweep_rows")}
    if not os.path.exists(os.path.join(src, names["bench"])):
        continue
    for k, nm in names.items():
        if os.path.exists(os.path.join(src, nm)):
            shutil.copy(os.path.join(src, nm), os.path.join(dst, nm.replace(tag, out, 1)))
    topo = f"{tag}topo_{n}gpu.txt"
    if os.path.exists(os.path.join(src, topo)):
        shutil.copy(os.path.join(src, topo), os.path.join(dst, topo.replace(tag, out, 1)))
    b = last_json(os.path.join(src, names["bench"]))
    row = {"n_gpus": n, "bench_step_ms": round(b["ms_per_step"], 4), "bench_audio_s_per_s": round(b["value"]),
           "clocks": b.get("clocks")}
    s = b.get("sustained")
    if s:
        row["sustained"] = {k: s[k] for k in ("seconds", "ms_per_step", "value", "sm_mhz_median", "power_w_median", "reasons") if k in s}
    e = b["e2e"]
    row["e2e_features_out"] = {"audio_s_per_s": round(e["value"]), "ms_per_call": round(e["ms_per_call"], 3),
                               "probe_ms": round(e["probe"]["ms"], 3),
                               "frac_of_probe_bound": round(e["probe"]["e2e_frac_of_probe_bound"], 3)}
    es = b.get("e2e_stats")
    if es:
        row["e2e_stats_pcm_to_moments"] = {"audio_s_per_s": round(es["value"]), "ms_per_call": round(es["ms_per_call"], 3),
                                           "probe_ms": round(es["probe_ms"], 3),
                                           "probe_bound_audio_s_per_s": round(es["probe_bound_audio_s_per_s"])}
    for k, key in (("corpus_dev", "corpus_device"), ("corpus_host", "corpus_host_pcm")):
        p = os.path.join(src, names[k])
        if os.path.exists(p):
            c = last_json(p)
            row[key] = {kk: c[kk] for kk in ("ms", "cold_ms", "audio_s_per_s", "shard_imbalance_max_over_mean",
                                            "sub_batches_rank0", "pipeline_chunks_rank0", "h2d_bytes_rank0") if kk in c}
    p = os.path.join(src, names["sweep_rows"])
    if os.path.exists(p):
        sw = json.load(open(p))
        row["sweep"] = [{"config": r["config"].split("per GPU, ")[-1], "ms": r["ms"], "audio_s_per_s": r["audio_s_per_s"]} for r in sw]
    rows.append(row)
base = next((r for r in rows if r["n_gpus"] == 1), None)
if base:
    for r in rows:
        n = r["n_gpus"]
        r["weak_scaling_efficiency_step"] = round(r["bench_audio_s_per_s"] / (n * base["bench_audio_s_per_s"]), 3)
        if "corpus_device" in r and "corpus_device" in base:
            r["strong_scaling_efficiency_corpus_device"] = round(base["corpus_device"]["ms"] / (n * r["corpus_device"]["ms"]), 3)
json.dump({"note": "tools/run_multi.sh N on one box (N = 1, 2, 4, 8 B200); bench.py, tools/run_corpus.py in both modes, "
                   "tools/run_configs.py --sweep-only; digest by tools/scaling_digest.py. The host-buffer legs are bounded by "
                   "the box (see e2e probe: bare cudaMemcpyAsync of the same bytes on all ranks at once).",
           "rows": rows}, open(os.path.join(dst, f"{out}scaling.json"), "w"), indent=1)
for r in rows:
    print(r["n_gpus"], r["bench_step_ms"], r["bench_audio_s_per_s"], r.get("weak_scaling_efficiency_step"),
          "| sust", round(r.get("sustained", {}).get("value", 0)), r.get("sustained", {}).get("sm_mhz_median"), r.get("sustained", {}).get("power_w_median"),
          "| e2e", r["e2e_features_out"], "| stats", r.get("e2e_stats_pcm_to_moments"),
          "| corpus", r.get("corpus_device", {}).get("ms"), r.get("strong_scaling_efficiency_corpus_device"), r.get("corpus_host_pcm", {}).get("ms"))
