"""Count the SASS instructions of one loop iteration of int_issue_peak_kernel (csrc/dev/peak.cu) and check the
constant HTN_INT_INSTR_PER_ITER the integer-issue roofline is computed from.  Run after a build (CPU only)."""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
obj = os.path.join(ROOT, "htnorm_b200", "csrc", "_obj", "peak.cu.o")
sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
body, on = [], False
for line in sass.splitlines():
    if "Function :" in line:
        on = "int_issue_peak_kernel" in line
    elif on:
        m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);", line)
        if m:
            body.append((int(m.group(1), 16), m.group(2).strip()))
back = [(a, t) for a, t in body if re.search(r"BRA 0x([0-9a-f]+)", t) and int(re.search(r"BRA 0x([0-9a-f]+)", t).group(1), 16) < a]
addr, text = back[0]
target = int(re.search(r"BRA 0x([0-9a-f]+)", text).group(1), 16)
n = sum(1 for a, _ in body if target <= a <= addr)
src = open(os.path.join(ROOT, "htnorm_b200", "csrc", "dev", "peak.cu")).read()
const = float(re.search(r"#define HTN_INT_INSTR_PER_ITER ([0-9.]+)", src).group(1))
print(f"loop body: {n} SASS instructions (0x{target:x}..0x{addr:x}); HTN_INT_INSTR_PER_ITER = {const}")
sys.exit(0 if n == int(const) else 1)
