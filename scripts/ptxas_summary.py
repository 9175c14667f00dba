"""Registers / spills / barriers per kernel from the -Xptxas -v logs the Makefile leaves next to the sources."""
import glob, os, re, subprocess
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "cross_attention_renderer_b200", "csrc")
print("# nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -Xptxas -v  (see csrc/Makefile)")
for log in sorted(glob.glob(os.path.join(root, "*.ptxas.log"))):
    txt = open(log).read()
    for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'.*?\n.*?\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers, used (\d+) barriers(?:, (\d+) bytes smem)?", txt):
        f = m.group(1)
        mang = f[f.index("_ZN"):] if "_ZN" in f else f
        name = subprocess.run(["c++filt", mang], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"car::\(anonymous namespace\)::", "", name)
        name = re.sub(r"\(.*", "", name)
        print(f"{os.path.basename(log)[:-10]:16s} {name[:48]:48s} regs={m.group(5):>3s} spill_st={m.group(3):>3s}B spill_ld={m.group(4):>3s}B stack={m.group(2):>3s}B barriers={m.group(6)} static_smem={m.group(7) or 0}B")
