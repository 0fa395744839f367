"""Writes the SASS evidence file (profiles/rNN_sass_excerpt.txt): per-kernel statistics of libvsiq.so and, for the hot
kernels, the address-ordered list of the instructions that matter -- 256-bit global accesses, bulk asynchronous copies
and mbarrier waits, programmatic-dependent-launch instructions, local-memory traffic, atomics, barriers -- so that "no
STL / LDL between the loads and the stores of the hot loop" can be read off.  Runs without a GPU.

    python tools/sass_excerpt.py > profiles/r02_sass_excerpt.txt
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "vsiquantization_b200", "libvsiq.so")
HOT = [  # (substring of the demangled name, why it is listed)
    ("fq_fwd_kernel<256, 8, false>", "UniformQuantizer forward, per tensor / per channel rows"),
    ("fq_bwd_ste_kernel<256, 8, false>", "STE backward -- the dominant kernel of the contract bench"),
    ("lsq_bwd_pt_kernel<8, 0, false, false>", "per-tensor LSQ backward (dx + dscale)"),
    ("lsq_bwd_pc_kernel<256, 8, true, true>", "per-channel LSQ backward, NCHW (dx + dscale[C] + dzp[C], ReLU fused)"),
    ("mt_bwd_kernel", "weight bank backward (all weight tensors of a step)"),
    ("ci_fwd_kernel<true, true, 1>", "NHWC epilogue forward: bias + ReLU + per-channel fake-quant"),
    ("ci_bwd_tma_kernel<true, true, 1, true>", "NHWC epilogue backward, inputs staged by bulk async copies"),
    ("ci_finalize_kernel", "combine kernel, a programmatic dependent of the streaming kernel"),
    ("ci_observe_kernel", "NHWC per-channel observer"),
    ("fq_codes_kernel<4, false>", "packed int4 code export"),
    ("ci_fwd2_kernel<true, true, 1, false>", "two-output epilogue: y and the next layer's quantize_inp tensor"),
    ("ci_epilogue_observe_kernel<1, 1>", "calibration epilogue: bias + ReLU + per-tensor observer in one pass"),
]
KEEP = re.compile(r"\b(LDG\S*|STG\S*|LDL\S*|STL\S*|UBLKCP\S*|SYNCS\S*|ACQBULK|PREEXIT|LDS\.128|ATOMG\S*|RED\S*|BAR\S*)\b")


def main():
    stats = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_stats.py"), "--all"], capture_output=True,
                           text=True, check=True).stdout
    print("# libvsiq.so (sm_100a) -- per-kernel SASS statistics (tools/sass_stats.py --all)")
    print("# LDG256 / STG256: 256-bit global accesses (LDG.E...256, sm_100 only); UBLKCP: bulk asynchronous copy (TMA, 1-D);")
    print("# SYNCS: mbarrier operations; STL / LDL: local-memory stores / loads (register spills or stack objects)\n")
    print(stats)
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    blocks, cur = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            blocks[cur] = []
        elif cur and re.search(r"^\s+/\*[0-9a-f]{4}\*/", line):
            blocks[cur].append(re.sub(r"\s+/\*\s*0x[0-9a-f]+\s*\*/\s*$", "", line).rstrip())
    names = subprocess.run(["c++filt"], input="\n".join(blocks), capture_output=True, text=True).stdout.splitlines()
    dem = dict(zip(names, blocks))
    for key, why in HOT:
        hit = [n for n in dem if key in n]
        if not hit:
            print(f"\n## {key}: not found")
            continue
        name = hit[0]
        body = blocks[dem[name]]
        print(f"\n## {re.sub(r'[(].*', '', name)}\n#  {why}; {len(body)} instructions; memory / synchronisation / call instructions in address order:")
        for line in body:
            if KEEP.search(line):
                print(line)


if __name__ == "__main__":
    main()
