"""The reference-compatible bench driver on the reference's own small fixture (gist_1000 x gist_test)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CONFIG = """label = "{label}"
dist = "L2Sqr"
gnd_path = "{d}/gnd.bin"
index_cache = "{d}/index.bin"
bench_output = "{d}/bench.toml"

[ef.range]
start = {ef0}
end = {ef1}
step = {step}

[algorithm.{algo}]
{algo_body}
{pq}
[base]
dim = 960
data_type = "float32"
data_path = "{d}/base.bin"

[test]
dim = 960
data_type = "float32"
data_path = "{d}/test.bin"
"""
IVF_BODY = "k = 16\nk_means_size = 300\nk_means_max_iter = 20\nk_means_tol = 1e-6\n"
PQ = ('[PQ]\npq_cache = "{d}/pq.bin"\ndist = "L2Sqr"\nn_bits = 4\nm = 240\nk_means_size = 300\n'
      'k_means_max_iter = 5\nk_means_tol = 1e-6\n')


def test_driver_runs_reference_style_configs(tmp_path, fixtures):
    from lab_1806_vec_db_b200 import formats as F
    d = str(tmp_path)
    F.save_raw(f"{d}/base.bin", fixtures["base"])
    F.save_raw(f"{d}/test.bin", fixtures["test"][:100])
    open(f"{d}/ivf.toml", "w").write(CONFIG.format(label="IVF", d=d, ef0=4, ef1=16, step=4, algo="IVF",
                                                   algo_body=IVF_BODY, pq=""))
    open(f"{d}/pq.toml", "w").write(CONFIG.format(label="Flat+PQ", d=d, ef0=100, ef1=200, step=100, algo="Flat",
                                                  algo_body="", pq=PQ.format(d=d)))
    for cfg in ("ivf.toml", "pq.toml"):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench_driver.py"), f"{d}/{cfg}"], capture_output=True,
                           text=True, cwd=ROOT, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    gnd = F.load_ground_truth(open(f"{d}/gnd.bin", "rb").read())
    assert len(gnd) == 100 and all(len(g) == 10 for g in gnd)
    res = F.load_result_list(open(f"{d}/bench.toml").read())
    by = {r["label"]: r for r in res["results"]}
    assert by["IVF"]["ef"] == [4, 8, 12, 16] and by["Flat+PQ"]["ef"] == [100, 200]
    assert by["IVF"]["recall"][-1] == 1.0            # all 16 lists probed = exact
    assert np.all(np.diff(by["IVF"]["recall"]) >= -1e-9)
    assert by["Flat+PQ"]["recall"][-1] > 0.9          # t_bench_1e4.toml reports 0.99+ for Flat+PQ at ef 100-200
    pq = F.load_pq_table(open(f"{d}/pq.bin", "rb").read())
    assert pq.m == 240 and pq.encoded_vec_set.shape == (1000, 120) and len(pq.group_k_means) == 240


def test_driver_runs_hnsw_configs(tmp_path, fixtures):
    """config/bench_hnsw.toml and config/bench_pq_240_hnsw.toml shapes (HNSW, HNSW + PQ) on the shipped fixture."""
    from lab_1806_vec_db_b200 import formats as F
    d = str(tmp_path)
    F.save_raw(f"{d}/base.bin", fixtures["base"])
    F.save_raw(f"{d}/test.bin", fixtures["test"][:100])
    body = "max_elements = 1000\nef_construction = 200\n"
    open(f"{d}/hnsw.toml", "w").write(CONFIG.format(label="HNSW", d=d, ef0=40, ef1=120, step=40, algo="HNSW", algo_body=body,
                                                    pq=""))
    open(f"{d}/hnsw_pq.toml", "w").write(CONFIG.format(label="HNSW+PQ m=240", d=d, ef0=100, ef1=300, step=100, algo="HNSW",
                                                       algo_body=body, pq=PQ.format(d=d)))
    for cfg in ("hnsw.toml", "hnsw_pq.toml"):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench_driver.py"), f"{d}/{cfg}"], capture_output=True,
                           text=True, cwd=ROOT, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    res = F.load_result_list(open(f"{d}/bench.toml").read())
    by = {r["label"]: r for r in res["results"]}
    assert by["HNSW"]["ef"] == [40, 80, 120] and by["HNSW+PQ m=240"]["ef"] == [100, 200, 300]
    assert by["HNSW"]["recall"][-1] >= 0.99          # the reference's 10k-row HNSW reaches 0.9927 at ef = 120
    assert by["HNSW+PQ m=240"]["recall"][-1] >= 0.97
