"""Build and time compile-time variants of the coder kernels (development aid).

    python tools/variants.py build  name1:-DFOO=1,-DBAR=2  name2:...      # here (nvcc cross-compiles)
    python tools/variants.py run [images]                                  # on the GPU box

`build` recompiles rans_encode.cu / rans_decode.cu / cdf_tables.cu with the extra flags and links
them with the other objects of the product build into tools/_build/var_<name>/libflic_b200.so
(travels to the GPU box).  `run` loads each variant in a fresh process (FLIC_B200_LIB), checks the
round trip on the sweep partition and prints device-timed encode / decode rates.
"""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "finalproject-losslessimagecompression_b200")
OUT = os.path.join(ROOT, "tools", "_build")
VAR_SRC = ["rans_encode.cu", "rans_decode.cu", "rans_decode_coop.cu", "cdf_tables.cu"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--fmad=false", "-Xcompiler", "-fPIC"]


def build(specs):
    sys.path.insert(0, PKG)
    import build as b
    b.build()
    procs = []
    for spec in specs:
        name, _, fl = spec.partition(":")
        d = os.path.join(OUT, "var_" + name)
        os.makedirs(d, exist_ok=True)
        extra = [f for f in fl.split(",") if f]
        for src in VAR_SRC:
            o = os.path.join(d, src.replace(".cu", ".o"))
            cmd = ["nvcc", *FLAGS, *extra, "-Xptxas=-v", "-I", os.path.join(PKG, "csrc"), "-I", os.path.join(ROOT, "include"),
                   "-c", os.path.join(PKG, "csrc", src), "-o", o]
            procs.append((name, src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for name, src, p in procs:
        out, _ = p.communicate()
        if p.returncode:
            print(out); raise SystemExit(f"{name}/{src} failed")
        lines = out.splitlines()
        for i, l in enumerate(lines):
            if "lane_kernelILi4" in l and "Compiling" in l:
                print(name, src, [x for x in lines[i:i + 4] if "Used" in x or "spill" in x])
    for spec in specs:
        name = spec.partition(":")[0]
        d = os.path.join(OUT, "var_" + name)
        objs = [os.path.join(d, s.replace(".cu", ".o")) for s in VAR_SRC]
        objs += [os.path.join(PKG, "_obj", s.replace(".cu", ".o")) for s in b.SOURCES if s not in VAR_SRC]
        subprocess.check_call(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
                               "-o", os.path.join(d, "libflic_b200.so"), *objs])
        print("built", name)


def run_one(images):
    sys.path.insert(0, ROOT)
    import torch
    from flic_b200 import rans
    n = images * 12288
    g = torch.Generator(device="cuda").manual_seed(1)
    mean = torch.randint(-256, 257, (n,), device="cuda", generator=g).float() / 256
    scale = torch.exp(10 * torch.rand(n, device="cuda", generator=g) - 5) / 256
    x = (torch.round((mean.double() + scale.double() * (10 * torch.rand(n, device="cuda", generator=g).double() - 5)) * 256) / 256).float()
    segs = torch.cat([torch.full((images,), sg, device="cuda") for sg in (6144, 3072, 3072)])
    off = torch.cat([torch.zeros(1, dtype=torch.int64, device="cuda"), torch.cumsum(segs, 0)])
    ws = rans.Workspace()
    out = torch.empty(n, device="cuda")
    enc = rans.encode_streams(x, mean, scale, off, workspace=ws, own_output=False)
    xr, end, st = rans.decode_streams(enc, mean, scale, off, out=out)
    ok = bool(torch.equal(xr, x)) and not bool(st.any()) and bool((end == (1 << 32)).all())
    csum = int(enc.words[: enc.n_words()].to(torch.int64).sum().item()) ^ int(enc.final_states.sum().item())

    def timed(fn, reps=5):
        fn(); torch.cuda.synchronize()
        best = 1e9
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        return best
    te = timed(lambda: rans.encode_streams(x, mean, scale, off, workspace=ws, own_output=False))
    td = timed(lambda: rans.decode_streams(enc, mean, scale, off, out=out))
    print(json.dumps({"ok": ok, "checksum": csum, "enc_Gsym": round(n / te / 1e6, 1), "dec_Gsym": round(n / td / 1e6, 1),
                      "enc_ms": round(te, 3), "dec_ms": round(td, 3)}))


def run(images):
    libs = [("product", os.path.join(PKG, "libflic_b200.so"))]
    for d in sorted(os.listdir(OUT)):
        p = os.path.join(OUT, d, "libflic_b200.so")
        if d.startswith("var_") and os.path.exists(p):
            libs.append((d[4:], p))
    for name, lib in libs:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "one", str(images)], capture_output=True, text=True,
                           env={**os.environ, "FLIC_B200_LIB": lib})
        print(name, (r.stdout.strip().splitlines() or [r.stderr[-300:]])[-1], flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2:])
    elif sys.argv[1] == "one":
        run_one(int(sys.argv[2]))
    else:
        run(int(sys.argv[2]) if len(sys.argv) > 2 else 131072)
