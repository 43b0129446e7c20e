"""Copy the text artefacts of the latest gpurun calls (gpurun_out/) into profiles/ under round-2 names (profiles/README.md)."""
import json, os, shutil, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PRO = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"


def last_json(path):
    for ln in reversed(open(path).read().splitlines()):
        if ln.startswith("{"):
            return json.loads(ln)
    return None


for src, dst in (("bench.log", "bench.json"), ("bench_reference.log", "bench_reference.json"), ("bench_n2.log", "bench_n2.json"),
                 ("bench_morld.log", "bench_morld.json"), ("bench_morld_serial.log", "bench_morld_serial.json"),
                 ("bench_morld_n2.log", "bench_morld_n2.json"), ("bench_n4.log", "bench_n4.json"), ("bench_morld_n4.log", "bench_morld_n4.json")):
    p = os.path.join(OUT, src)
    if os.path.exists(p):
        j = last_json(p)
        if j:
            json.dump(j, open(os.path.join(PRO, f"{tag}_{dst}"), "w"), indent=1)
            print("wrote", f"{tag}_{dst}")
for src, dst in (("bench_ab.log", "bench_ab.txt"), ("gemm_stats.log", "gemm_stats.txt"), ("golden_diag.log", "golden_diag.txt"),
                 ("sanitize_memcheck.log", "sanitize_memcheck.txt"), ("sanitize_racecheck.log", "sanitize_racecheck.txt"),
                 ("sanitize_synccheck.log", "sanitize_synccheck.txt"), ("sanitize_initcheck.log", "sanitize_initcheck.txt"),
                 ("sanitize_new_memcheck.log", "sanitize_new_memcheck.txt"), ("sanitize_new_racecheck.log", "sanitize_new_racecheck.txt"),
                 ("sanitize_new_synccheck.log", "sanitize_new_synccheck.txt"), ("qhead_time.log", "qhead_time.txt"),
                 ("pytest_gpu.log", "pytest_gpu.txt"), ("smoke.log", "smoke.txt")):
    p = os.path.join(OUT, src)
    if os.path.exists(p):
        shutil.copyfile(p, os.path.join(PRO, f"{tag}_{dst}"))
        print("wrote", f"{tag}_{dst}")
a, b = os.path.join(OUT, "gemm_error_single.log"), os.path.join(OUT, "gemm_error_split.log")
if os.path.exists(a) and os.path.exists(b):
    open(os.path.join(PRO, f"{tag}_gemm_error.txt"), "w").write(open(a).read() + "\n" + open(b).read())
    print("wrote", f"{tag}_gemm_error.txt")
