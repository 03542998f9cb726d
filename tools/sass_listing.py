#!/usr/bin/env python
"""Register / spill table (ptxas -v) and SASS opcode histograms of the hot kernels, for profiles/.

    python tools/sass_listing.py profiles/r02_sass_listing.md

Compiles the sources exactly as mimic_mda_b200/build.py does (sm_100a, -lineinfo) with -Xptxas -v, disassembles the
objects with cuobjdump -sass, and writes per kernel: registers, spills, shared memory, and the count of every opcode
family -- the mnemonics that prove which path a kernel takes (DFMA / DMMA for the fp64 pipes, UBLKCP for bulk TMA
copies, SYNCS for mbarriers, USETMAXREG for the register re-allocation of the warp-specialised stepper).
"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mimic_mda_b200 import build  # noqa: E402

HOT = [  # (source, demangled-name fragment, label)
    ("ensemble_qws_kernel.cu", "ensemble_qws_kernel<(int)9, (bool)1, (bool)0>", "ensemble_qws_kernel<9, T in registers> (C4 default: one CTA per SM)"),
    ("ensemble_qws_kernel.cu", "ensemble_qws_kernel<(int)9, (bool)0, (bool)0>", "ensemble_qws_kernel<9, T in shared memory> (two CTAs per SM)"),
    ("ensemble_quad_kernel.cu", "ensemble_quad_kernel<(int)9, (int)1>", "ensemble_quad_kernel<9,1>"),
    ("ensemble_quad_kernel.cu", "ensemble_quad_kernel<(int)9, (int)3>", "ensemble_quad_kernel<9,3>"),
    ("ensemble_quad_kernel.cu", "ensemble_quad_kernel<(int)9, (int)4>", "ensemble_quad_kernel<9,4>"),
    ("ensemble_ws_kernel.cu", "ensemble_ws_kernel<(int)9, (int)4", "ensemble_ws_kernel<9,4> (direct evaluation, fp64 DMMA)"),
    ("ensemble_kernel.cu", "ensemble_kernel<(int)9, (int)2", "ensemble_kernel<9,2> (reference-order stepper of ill-conditioned stores)"),
    ("lnpost_kernel.cu", "lnpost_tile_kernel<(int)9, (int)2>", "lnpost_tile_kernel<9,2> (batched entry point)"),
    ("lnpost_split_kernel.cu", "lnpost_split_kernel<(int)9>", "lnpost_split_kernel<9> (opt-in)"),
    ("lnpost_mma_kernel.cu", "lnpost_mma_kernel<(int)9>", "lnpost_mma_kernel<9>"),
]
WATCH = ["DFMA", "DMUL", "DADD", "DSETP", "DMMA", "UBLKCP", "SYNCS", "USETMAXREG", "LDS", "STS", "LDG", "STG", "LDGSTS", "BAR", "IMAD", "LOP3",
         "MUFU", "I2F", "SHFL", "ATOM", "RED", "LDL", "STL", "UTCMMA", "LDTM"]


def run(cmd):
    return subprocess.run(cmd, capture_output=True, text=True, check=True)


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_listing.md")
    nvcc = build.nvcc_path()
    flags = [f for f in build.NVCC_FLAGS if f not in ("-shared", "-cudart", "static")]
    lines = ["# SASS / ptxas listing of the hot kernels (sm_100a, nvcc %s)\n" % run([nvcc, "--version"]).stdout.strip().split("\n")[-2].strip(),
             "Made by `python tools/sass_listing.py`: `nvcc %s -Xptxas -v -c <source>` + `cuobjdump -sass`.\n" % " ".join(flags),
             "No `UTC*MMA` / `LDTM` anywhere: tcgen05 has no fp64 kind (north_star rules tensor cores out for this arithmetic); the fp64 tensor",
             "path that exists, `DMMA.8x8x4`, is used by the direct-evaluation steppers.  `UBLKCP` = bulk-TMA copies (`cp.async.bulk`),",
             "`SYNCS` = mbarrier operations, `USETMAXREG` = `setmaxnreg` of the warp-specialised stepper.\n"]
    by_src = collections.defaultdict(list)
    for src, frag, label in HOT:
        by_src[src].append((frag, label))
    with tempfile.TemporaryDirectory() as tmp:
        for src, wanted in by_src.items():
            obj = os.path.join(tmp, src + ".o")
            res = run([nvcc] + flags + ["-Xptxas", "-v", "-c", os.path.join(build.CSRC, src), "-o", obj])
            ptxas = res.stderr
            sass = run(["cuobjdump", "-sass", obj]).stdout
            # split the disassembly per function
            funcs = re.split(r"\n\s*Function : ", sass)[1:]
            names = {}
            for f in funcs:
                mangled = f.split("\n", 1)[0].strip()
                names[mangled] = f
            dem = subprocess.run(["cu++filt"] + list(names), capture_output=True, text=True).stdout.strip().split("\n")
            table = dict(zip(dem, names))
            for frag, label in wanted:
                hit = [d for d in table if frag in d]
                if not hit:
                    lines.append("## %s\n\nnot found in %s\n" % (label, src))
                    continue
                mangled = table[hit[0]]
                body = names[mangled]
                info = re.search(re.escape(mangled) + r"'.*?\n(.*?)\nptxas info\s*:\s*Used (\d+) registers(.*?)\n", ptxas, flags=re.S)
                spill = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", info.group(1)) if info else None
                ops = collections.Counter()
                n_inst = 0
                for ln in body.split("\n"):
                    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
                    if m:
                        n_inst += 1
                        ops[m.group(1).split(".")[0]] += 1
                lines.append("## %s\n" % label)
                lines.append("`%s`\n" % hit[0])
                if info:
                    lines.append("- registers per thread at launch: **%s**%s" % (info.group(2), info.group(3).strip().rstrip(",")))
                if spill:
                    lines.append("- stack frame %s B, spill stores %s B, spill loads %s B" % spill.groups())
                lines.append("- SASS instructions: %d" % n_inst)
                lines.append("- opcode families: " + ", ".join("%s %d" % (k, ops[k]) for k in WATCH if ops.get(k)))
                others = sorted(((v, k) for k, v in ops.items() if k not in WATCH), reverse=True)[:8]
                lines.append("- others (top): " + ", ".join("%s %d" % (k, v) for v, k in others) + "\n")
    with open(out_path, "w") as fh:
        fh.write("\n".join(lines) + "\n")
    print(out_path)


if __name__ == "__main__":
    main()
