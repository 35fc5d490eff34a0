"""Static SASS opcode histogram of selected kernels of an object file (cuobjdump -sass) -> markdown table.

    python scripts/sass_histogram.py path-tracing__ray-tracer_b200/build/rt_f32.o 'shade_kernel<float, b2rt::PcgRng, 3>' ...
"""
import collections, re, subprocess, sys
obj, wanted = sys.argv[1], sys.argv[2:]
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
name, kernels = None, collections.OrderedDict()
for l in out.split("\n"):
    m = re.search(r"Function : (\S+)", l)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kernels[name] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_]*)", l)
    if m and name:
        kernels[name][m.group(1)] += 1
PIPE = {"FFMA": "fma", "FMUL": "fma", "FADD": "fma", "IMAD": "fma", "HFMA2": "fma", "FFMA32I": "fma", "FMUL32I": "fma", "FADD32I": "fma",
        "LOP3": "alu", "FSETP": "alu", "ISETP": "alu", "SEL": "alu", "FSEL": "alu", "FMNMX": "alu", "FMNMX3": "alu", "IADD3": "alu",
        "SHF": "alu", "PRMT": "alu", "LEA": "alu", "PLOP3": "alu", "MOV": "alu", "VIADD": "alu", "IMNMX": "alu", "VIMNMX": "alu",
        "POPC": "alu", "FLO": "alu", "I2FP": "alu", "F2FP": "alu", "VIMNMX3": "alu",
        "MUFU": "xu", "I2F": "xu", "F2I": "xu", "F2F": "xu",
        "LDS": "lsu", "STS": "lsu", "LDG": "lsu", "STG": "lsu", "LDL": "lsu", "STL": "lsu", "ATOMG": "lsu", "ATOMS": "lsu", "RED": "lsu",
        "LDGSTS": "lsu", "LDC": "lsu/const", "LDCU": "uniform", "BRA": "cbu", "BSSY": "cbu", "BSYNC": "cbu", "EXIT": "cbu", "CALL": "cbu",
        "RET": "cbu", "WARPSYNC": "cbu", "VOTE": "alu", "SHFL": "lsu", "S2R": "misc", "CS2R": "misc", "NOP": "misc", "BAR": "cbu",
        "DEPBAR": "misc", "LDGDEPBAR": "misc"}
for w in wanted:
    for k, c in kernels.items():
        if w in k:
            tot = sum(c.values())
            print(f"### `{k[:100]}`\n\n{tot} SASS instructions ({tot * 16} B)\n")
            print("| opcode | count | % | pipe |\n|---|---|---|---|")
            for op, n in c.most_common(28):
                print(f"| {op} | {n} | {100 * n / tot:.1f} | {PIPE.get(op, '?')} |")
            by = collections.Counter()
            for op, n in c.items():
                by[PIPE.get(op, "other")] += n
            print("\nby pipe: " + ", ".join(f"{p} {100 * n / tot:.1f} %" for p, n in by.most_common()) + "\n")
