"""Static look at the hot loop of a kernel in an object file: instruction mix, FFMA / FMNMX source operands that are not
served by the operand-reuse cache, and a register-port cost under the two-banks-by-parity model of tools/ubench_banks.cu.

    python tools/sass_bank_score.py dicp_b200/lib/nn_kernels.o nn_scan_f32_kernelILi8

Used in round 2 to pre-screen 48 (loop order x interleave x lane rotation x CTAs per SM) variants of the NN filter loop;
the model ranks them, the GPU does not agree with the ranking (profiles/r02_nn_notes.md) -- kept as a SASS reading aid."""
import re, subprocess, sys
def kernel_sass(obj, name_part):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    blocks = out.split("Function : ")
    for b in blocks[1:]:
        if name_part in b.split("\n", 1)[0]:
            return b
    return None
def parse(b):
    ins = []
    for l in b.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    return ins
def main_loop(ins):
    best = None
    addr_index = {a: i for i, (a, _) in enumerate(ins)}
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < a and tgt in addr_index:
                body = ins[addr_index[tgt]:i + 1]
                nf = sum(1 for _, x in body if re.match(r"(@!?U?P\d+\s+)?FFMA", x))
                if best is None or nf > best[0]:
                    best = (nf, body)
    return best[1] if best else []
def score(body):
    prev = None
    stats = {}
    port = 0
    n = 0
    for _, t in body:
        m = re.match(r"(@!?U?P\d+\s+)?(\S+)\s+(.*)", t)
        if not m:
            prev = None; continue
        op = m.group(2); args = [a.strip() for a in m.group(3).split(",")]
        n += 1
        if op.startswith("FFMA") or op.startswith("FMNMX") or op.startswith("FSEL") or op.startswith("FSETP") or op.startswith("SEL"):
            srcs = args[1:4]
            if op.startswith("FSETP"): srcs = args[2:4]
            fresh = []
            for i, s in enumerate(srcs):
                reg = s.replace(".reuse", "").replace("-", "").replace("|", "")
                if not re.match(r"R\d+$", reg): continue
                hit = prev is not None and i < len(prev) and prev[i] == (reg, True)
                if not hit: fresh.append(int(reg[1:]))
            par = [r % 2 for r in fresh]
            c = max(par.count(0), par.count(1), 1)
            port += c
            key = (op.split(".")[0], len(fresh), c)
            stats[key] = stats.get(key, 0) + 1
            prev = [(s.replace(".reuse", "").replace("-", "").replace("|", ""), ".reuse" in s) for s in srcs]
        else:
            port += 1
            prev = None
    return n, port, stats
if __name__ == "__main__":
    b = kernel_sass(sys.argv[1], sys.argv[2])
    body = main_loop(parse(b))
    n, port, stats = score(body)
    nf = sum(v for k, v in stats.items() if k[0] == "FFMA")
    print("%s: loop %d instr, %d FFMA, port-cycles %d (%.3f per FFMA-eval-equivalent), 3-fresh FFMA %d" % (
        sys.argv[1], n, nf, port, port / (nf / 3.0), sum(v for k, v in stats.items() if k[0] == "FFMA" and k[1] == 3)))
