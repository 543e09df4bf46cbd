"""Build / run configuration variants of the fused kernels.

    python tools/tune.py build tune/spec.json      # here (no GPU): one libyf per variant, in parallel
    python tools/tune.py run [res] [batch]         # on the GPU box: profile every tune/lib_*.so

spec.json: {"tag": {"YF_CFGRES2": "IrbCfg<...>", ...}, ...}
"""
import json
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "yolo_fastest_b200", "csrc")
TUNE = os.path.join(ROOT, "tune")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared"]


def build_one(item):
    tag, defs = item
    out = os.path.join(TUNE, "lib_%s.so" % tag)
    hdr = os.path.join(TUNE, "cfg_%s.h" % tag)          # nvcc splits -D values at commas, so the overrides go through a pre-include
    with open(hdr, "w") as f:
        f.write("".join("#define %s %s\n" % (k, v) for k, v in defs.items()))
    cmd = ["/usr/local/cuda/bin/nvcc"] + FLAGS + ["--pre-include", hdr, "-o", out, os.path.join(CSRC, "yf_api.cu")]
    r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    return tag, r.returncode, r.stderr[-2000:]


def main():
    if sys.argv[1] == "build":
        spec = json.load(open(sys.argv[2]))
        for f in os.listdir(TUNE):
            if f.startswith("lib_") and f.endswith(".so"):
                os.remove(os.path.join(TUNE, f))
        with ThreadPoolExecutor(8) as ex:
            for tag, rc, err in ex.map(build_one, spec.items()):
                print(tag, "ok" if rc == 0 else "FAILED\n" + err)
    else:
        res = sys.argv[2] if len(sys.argv) > 2 else "512x640"
        batch = sys.argv[3] if len(sys.argv) > 3 else "256"
        libs = [None] + sorted(os.path.join(TUNE, f) for f in os.listdir(TUNE) if f.startswith("lib_") and f.endswith(".so"))
        rows = []
        for lib in libs:
            env = dict(os.environ)
            if lib:
                env["YF_B200_LIB"] = lib
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "profile_groups.py"), res, batch], env=env,
                               capture_output=True, text=True, timeout=600)
            line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ""
            try:
                rows.append(json.loads(line))
            except Exception:
                print("FAILED", lib, r.stderr[-1500:])
        base = rows[0]["groups"] if rows else {}
        for row in rows:
            tag = os.path.basename(row["lib"])
            diff = {k: v for k, v in row["groups"].items() if abs(v - base.get(k, 0)) > 0.03 * base.get(k, 1)}
            print("%-28s total %.3f ms  err %.2e  %s" % (tag, row["total_ms"], row["err"], json.dumps(diff if row is not rows[0] else row["groups"])))


if __name__ == "__main__":
    main()
