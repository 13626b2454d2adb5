"""Upper bounds for the tile kernel (C4, FP64, tile-ordered): build tri_tile9.cu with -DHIDENN_ABL=k (k = 1, 2, 3: two
thirds / one third / none of the per-element shared-memory exchange; results wrong on purpose) into gpurun_alt/ and time
the kernel alone.  `--build` here (no GPU), `--run` on the GPU box.  Extra -D flags: --defs "-DX=1 -DY=2" --tag name."""
import argparse, os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ALT = os.path.join(ROOT, "gpurun_alt")


def build(tag, defs, srcs=("tri_tile9.cu",)):
    from hidenn_fem_b200 import build as b
    b.build()
    os.makedirs(ALT, exist_ok=True)
    objs = []
    for src in b.sources():
        base = os.path.basename(src)
        obj = os.path.join(os.path.dirname(b.LIB), "build", base[:-3] + ".o")
        if base in srcs:
            obj = os.path.join(ALT, f"{tag}_{base[:-3]}.o")
            subprocess.check_call([b._nvcc(), *b.NVCC_FLAGS, *defs, "-c", src, "-o", obj])
        objs.append(obj)
    out = os.path.join(ALT, f"lib_{tag}.so")
    subprocess.check_call([b._nvcc(), "-shared", *b.NVCC_FLAGS, "-o", out, *objs])
    return out


def run_one():
    import torch, bench
    args = bench.parse()
    dev = torch.device("cuda:0")
    if os.environ.get("ABLATE_ROTATE"):      # every element's corners rotated at random: all nine pair classes occur (gmsh-like)
        import numpy as np
        from hidenn_fem_b200 import meshgen
        orig = meshgen.plate_mesh

        def rotated(*a, **k):
            mm = orig(*a, **k)
            rot = np.random.default_rng(0).integers(0, 3, mm.connectivity.shape[0])
            mm.connectivity = np.take_along_axis(mm.connectivity, (np.arange(3)[None, :] + rot[:, None]) % 3, axis=1)
            return mm
        meshgen.plate_mesh = rotated
    m, model, loss_fn, _ = bench.make_workload(args, 0, 1, dev, torch.float64, "tiles", args.elems, args.tile_nodes)
    ms = min(bench.time_kernel(model, loss_fn, 20, 5) for _ in range(3))
    info = model._plan().info
    rec = {"lib": os.path.basename(os.environ.get("HIDENN_LIB", "default")), "kernel_us": round(ms * 1e3, 1),
           "pairs": os.environ.get("HIDENN_PLAN_PAIRS", "0"), "tile_nodes": args.tile_nodes, "n_tiles": info["n_tiles"],
           "entries_per_tile": round(info.get("pair_entries", 0) / info["n_tiles"], 1), "n_pairs": info.get("n_pairs", 0)}
    import re
    if "prof" in rec["lib"]:      # HIDENN_PROF9 build: per-warp wait cycles sit in the tile-energy scratch
        torch.cuda.synchronize()
        plan = model._plan()
        nw = int((re.search(r"_w(\d+)", rec["lib"]) or [0, "24"])[1])
        sc = loss_fn._scratch(plan, dev, torch.float64)[:148 * nw * 4].reshape(148, nw, 4).cpu().numpy()
        ew = int((re.search(r"_e(\d+)", rec["lib"]) or [0, "12"])[1])      # tags: prof_e12_l2 ...
        lw = int((re.search(r"_l(\d+)", rec["lib"]) or [0, "2"])[1])
        # paired builds: 16 element warps; of the 8 others, warps 20 and 21 load (HIDENN_WS_LPOS = 4)
        if "_pe" in rec["lib"]:
            roles = (("element_active", list(range(0, 13))), ("element_idle", [14, 15]), ("fold", [16, 17, 18, 19, 22, 23]), ("loader", [20, 21]))
        else:
            roles = (("element", list(range(0, ew))), ("fold", list(range(ew, nw - lw))), ("loader", list(range(nw - lw, nw))))
        for name, sl in roles:
            w = sc[:, sl]
            tot = w[..., 0].mean()
            rec[name] = {"total_cyc": round(float(tot)), "waitA_frac": round(float((w[..., 1] / w[..., 0]).mean()), 3),
                         "waitB_frac": round(float((w[..., 2] / w[..., 0]).mean()), 3), "tiles": float(w[..., 3].mean())}
    print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--build", action="store_true")
    ap.add_argument("--run", action="store_true")
    ap.add_argument("--one", action="store_true")
    ap.add_argument("--defs", default="")
    ap.add_argument("--tag", default="")
    a, rest = ap.parse_known_args()
    if a.one:
        sys.argv = [sys.argv[0]] + rest
        run_one()
    elif a.build:
        if a.tag:
            print(build(a.tag, a.defs.split()))
        else:
            for k in (1, 2, 3):
                print(build(f"abl{k}", [f"-DHIDENN_ABL={k}"]))
    elif a.run:
        libs = [None] + sorted(os.path.join(ALT, f) for f in os.listdir(ALT) if f.startswith("lib_") and f.endswith(".so"))
        for lib in libs:
            env = dict(os.environ)
            if lib:
                env["HIDENN_LIB"] = lib
            subprocess.call([sys.executable, os.path.abspath(__file__), "--one"], env=env, cwd=ROOT)
