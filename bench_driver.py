#!/usr/bin/env python
"""bench_driver.py — runs a reference BenchConfig TOML (examples/bench.rs:70-92) on the GPU backend and appends a
`[[results]]` block in the reference's ResultList TOML format (bench.rs:312-368), so curves are directly comparable
with data/t_bench*.toml. Supported `algorithm` tables: `HNSW`, `IVF` and `Flat`, each with or without `[PQ]`
(HNSW + PQ = IndexPQ::knn_pq on the graph, hnsw_index.rs:672-697).

  python bench_driver.py config/bench_10000_ivf.toml [--repeat-times N]

search_time is elapsed / (repeats * n_queries) in ms for the whole query batch (the reference's `-t` protocol:
aggregate throughput, bench.rs:414-425); recall is recall@10 against the Flat ground truth file (`gnd_path`, bincode
GroundTruth), which is generated with the exact GPU Flat scan when missing (src/bin/gen_gnd.rs protocol, k = 10).
"""
import argparse
import os
import time

import numpy as np

import lab_1806_vec_db_b200 as V
from lab_1806_vec_db_b200 import formats as F

DT = {"float32": np.float32, "uint8": np.uint8}


def load_set(cfg):
    return F.load_raw(cfg["data_path"], cfg["dim"], DT[cfg.get("data_type", "float32")], cfg.get("limit"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("bench_config_path")
    ap.add_argument("-r", "--repeat-times", type=int, default=1)
    ap.add_argument("--seed", type=int, default=42)  # bench.rs:371
    args = ap.parse_args()
    cfg = F.load_bench_config(args.bench_config_path)
    dist = cfg["dist"]
    base, test = load_set(cfg["base"]), load_set(cfg["test"])
    rng = np.random.default_rng(args.seed)
    vs = V.DeviceVecSet(base, dist)
    flat = V.FlatIndex(vs)
    k = 10
    if os.path.exists(cfg["gnd_path"]):
        gnd = F.load_ground_truth(open(cfg["gnd_path"], "rb").read())
    else:
        ids, _, cnt = flat.knn_batch(test, k)
        gnd = [ids[i, :cnt[i]] for i in range(len(test))]
        os.makedirs(os.path.dirname(cfg["gnd_path"]) or ".", exist_ok=True)
        open(cfg["gnd_path"], "wb").write(F.dump_ground_truth(gnd))
    algo = cfg["algorithm"]
    pq = None
    if cfg.get("PQ"):
        p = cfg["PQ"]
        pq = V.PQTable.from_vec_set(vs, base, V.PQConfig(p["n_bits"], p["m"], F.DIST_TOML[p["dist"]], p.get("k_means_size"),
                                                         p["k_means_max_iter"], p["k_means_tol"]), rng)
        if p.get("pq_cache"):
            os.makedirs(os.path.dirname(p["pq_cache"]) or ".", exist_ok=True)
            open(p["pq_cache"], "wb").write(F.dump_pq_table(F.pq_table_record(pq), base.dtype))
    if "IVF" in algo:
        a = algo["IVF"]
        index = V.IVFIndex.from_vec_set(vs, base, dist, V.IVFConfig(a["k"], a.get("k_means_size"), a["k_means_max_iter"],
                                                                    a["k_means_tol"]), rng)
        search = lambda ef: index.knn_with_ef_batch(test, k, ef)  # noqa: E731
    elif "HNSW" in algo:
        a = algo["HNSW"]
        t0 = time.perf_counter()
        index = V.HNSWIndex(vs, V.HNSWConfig(a.get("max_elements", 0), a.get("ef_construction", 200), a.get("M", 16)), rng)
        print(f"HNSW build: {time.perf_counter() - t0:.2f} s for {len(base)} vectors", flush=True)
        if pq is None:
            search = lambda ef: index.knn_with_ef_batch(test, k, ef)  # noqa: E731
        else:
            search = lambda ef: index.knn_pq_batch(test, k, ef, pq)  # noqa: E731
    elif "Flat" in algo or "flat" in algo:
        if pq is None:
            search = lambda ef: flat.knn_batch(test, k)  # noqa: E731
        else:
            search = lambda ef: flat.knn_pq_batch(test, k, ef, pq)  # noqa: E731
    else:
        raise SystemExit(f"unsupported algorithm table {list(algo)}")
    res = {"label": cfg["label"], "ef": [], "search_time": [], "recall": []}
    for ef in cfg["ef_values"]:
        search(ef)  # warm-up
        t0 = time.perf_counter()
        for _ in range(args.repeat_times):
            ids, _, cnt = search(ef)
        ms = (time.perf_counter() - t0) * 1e3 / (args.repeat_times * len(test))
        rec = float(np.mean([F.recall(gnd[i], ids[i, :cnt[i]]) for i in range(len(test))]))
        res["ef"].append(ef)
        res["search_time"].append(ms)
        res["recall"].append(rec)
        print(f"ef={ef}: {ms:.6f} ms/query ({1e3 / ms:.0f} QPS), recall@10 {rec:.4f}", flush=True)
    out = cfg["bench_output"]
    old = F.load_result_list(open(out).read()) if os.path.exists(out) else {"title": "", "results": []}
    results = [r for r in old["results"] if r["label"] != res["label"]] + [res]
    title = old["title"] or f"Bench (N={len(base)}, dim={base.shape[1]}, B200 backend)"
    os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
    open(out, "w").write(F.dump_result_list(title, results))
    print("wrote", out)


if __name__ == "__main__":
    main()
