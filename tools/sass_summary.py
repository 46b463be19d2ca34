#!/usr/bin/env python3
"""SASS evidence for profiles/: instruction mix of the production kernels from `cuobjdump -sass` of the built library,
plus the full SASS of the headline kernel.  Runs without a GPU.

    python tools/sass_summary.py  ->  profiles/r2_sass_summary.md, profiles/r2_sass_k1w_1024b_mag_hann.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "kopek_b200", "lib", "libkopek_cuda.so")

# (label, regex on the demangled-ish function line)
KERNELS = [
    ("K1W-b 1024 |X| Hann (headline)", r"k1w_1024b_kernelILi1ELb1ELi4ELi3ELb0ELb0ELb0ELb1E"),
    ("K1W-b 1024 |X| Hann, FULL rows (1024 bins)", r"k1w_1024b_kernelILi1ELb1ELi4ELi3ELb0ELb0ELb1ELb1E"),
    ("K1W-b 1024 |X| Hann, 32-bit loads (odd hop)", r"k1w_1024b_kernelILi1ELb1ELi4ELi3ELb0ELb0ELb0ELb0E"),
    ("K1W-b 1024 |X| Hann, bulk-store epilogue", r"k1w_1024b_kernelILi1ELb1ELi4ELi3ELb1ELb0ELb0ELb1E"),
    ("K1W-256 |X| rect (configs[2])", r"k1w_256_kernelILi1ELb0ELi4ELi3ELb0ELb1E"),
    ("K1W-512 |X| Hann", r"k1w_512_kernelILi1ELb1ELi4ELi3ELb0ELb1E"),
    ("K1M 2048 dB Hann (configs[1])", r"k1m_kernelILi4ELi3ELi2ELi6ELi0ELb0ELb1E"),
    ("K1M 4096 |X| Hann (configs[4])", r"k1m_kernelILi8ELi1ELi2ELi3ELi0ELb0ELb1E"),
    ("K1W-b 1024 log bands Hann", r"k1w_1024b_kernelILi4ELb1ELi4ELi3ELb0ELb0ELb0ELb1E"),
    ("pcm16 -> f32 (stereo, vector path)", r"pcm16x2_to_f32_kernel"),
    ("detect_bpm fold (warp per sum)", r"bpm_fold_kernel"),
    ("K2 f64 drop-in, chunk", r"k2_c2c_f64_chunk"),
    ("K3 pass 1, N1 = 4096", r"k3_pass1ILi4096E"),
    ("K3 pass 2, N2 = 4096 (radix 64 x 64, production)", r"k3_pass2_r64"),
    ("synth_frames", r"synth_frames_kernel"),
]
GROUPS = collections.OrderedDict([
    ("FFMA", r"^FFMA"), ("FMUL", r"^FMUL"), ("FADD", r"^FADD(?!2)"), ("FADD2 (packed)", r"^FADD2"), ("MUFU", r"^MUFU"), ("DFMA/DMUL/DADD", r"^D(FMA|MUL|ADD)"),
    ("LDG", r"^LDG"), ("STG", r"^STG"), ("LDS", r"^LDS"), ("STS", r"^STS"), ("SHFL", r"^SHFL"),
    ("UBLKCP (bulk copy)", r"^UBLKCP"), ("UTMALDG/UTMASTG (tensor TMA)", r"^UTMA"), ("SYNCS (mbarrier)", r"^SYNCS"),
    ("BAR", r"^BAR"), ("WARPSYNC", r"^WARPSYNC"), ("SEL/FSEL", r"^F?SEL"), ("IMAD/IADD/LEA/LOP", r"^(IMAD|IADD|LEA|LOP|SHF|UIADD|UIMAD|ULEA)"),
    ("BRA", r"^BRA"),
])


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs = {}
    name = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
        elif name is not None:
            m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(.*?);", line)
            if m:
                ins = re.sub(r"^@!?U?P\d+\s+", "", m.group(1).strip())
                funcs[name].append(ins)
    out = ["# SASS instruction mix of the production kernels (cuobjdump -sass kopek_b200/lib/libkopek_cuda.so, sm_100a)", "",
           "Static counts over the whole kernel (prologue + frame loop + tail); the frame loops are fully unrolled, so the",
           "loop body dominates.  `UBLKCP` = `cp.async.bulk` (1-D TMA), `UTMALDG` = `cp.async.bulk.tensor` (2-D TMA),",
           "`SYNCS` = mbarrier operations.  No `HMMA`/`UTCMMA`: the path is HBM-bound f32 streaming work, tensor cores unused.", ""]
    hdr = "| kernel | total | " + " | ".join(GROUPS) + " |"
    out += [hdr, "|---|---:|" + "---:|" * len(GROUPS)]
    head_name = None
    for label, pat in KERNELS:
        hits = [f for f in funcs if re.search(pat, f)]
        if not hits:
            out.append(f"| {label} | not found |" + " |" * len(GROUPS))
            continue
        f = hits[0]
        if head_name is None:
            head_name = f
        ins = funcs[f]
        counts = [sum(1 for i in ins if re.match(rx, i)) for rx in GROUPS.values()]
        out.append(f"| {label} | {len(ins)} | " + " | ".join(str(c) for c in counts) + " |")
    tensor = sum(1 for f in funcs.values() for i in f if re.match(r"^(HMMA|UTCMMA|UTCHMMA|QGMMA|IMMA)", i))
    out += ["", f"Tensor-core instructions in the whole library: {tensor}.  Functions in the library: {len(funcs)}."]
    with open(os.path.join(ROOT, "profiles", "r2_sass_summary.md"), "w") as fh:
        fh.write("\n".join(out) + "\n")
    with open(os.path.join(ROOT, "profiles", "r2_sass_k1w_1024b_mag_hann.txt"), "w") as fh:
        fh.write(f"// cuobjdump -sass, function {head_name}\n" + "\n".join(funcs[head_name]) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    sys.exit(main())
