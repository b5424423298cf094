#!/usr/bin/env python3
"""ptx2expr -- print the floating-point expression tree behind every global store
(and every fp compare) of one PTX kernel.

Development tool (not part of the product or the oracle).  It is used to read
off exactly how nvcc contracted mul+add into fma in a kernel, so that the CPU
oracle (oracle/fdgs_oracle.c) and the CUDA kernels (explicit __fmaf_rn /
__fmul_rn / __fadd_rn) can reproduce the arithmetic bit for bit.

usage: ptx2expr.py file.ptx <kernel-name-substring> [--depth N]
"""
import re
import sys


def parse_kernel(text, name_sub):
    # split into entries
    entries = re.split(r"(?=\.visible \.entry|\.entry)", text)
    for e in entries:
        head = e.split("(", 1)[0]
        if name_sub in head and ".entry" in head:
            # cut at the closing brace of the entry body
            out = []
            for l in e.splitlines():
                out.append(l)
                if l.strip() == "}":
                    break
            e = "\n".join(out)
            # shorten parameter names
            e = re.sub(r"_Z\w+_param_(\d+)", r"param\1", e)
            return e
    raise SystemExit("kernel not found: " + name_sub)


FLOAT_IMM = re.compile(r"^0[fF]([0-9A-Fa-f]{8})$")
DBL_IMM = re.compile(r"^0[dD]([0-9A-Fa-f]{16})$")


def imm(tok):
    import struct
    m = FLOAT_IMM.match(tok)
    if m:
        v = struct.unpack(">f", bytes.fromhex(m.group(1)))[0]
        return "%.9gf" % v
    m = DBL_IMM.match(tok)
    if m:
        v = struct.unpack(">d", bytes.fromhex(m.group(1)))[0]
        return "%.17g" % v
    return tok


MAXLEN = 70
PARAM_NAMES = {}


class Sym:
    def __init__(self):
        self.defs = {}      # reg -> expr string
        self.multi = set()  # regs defined more than once
        self.count = {}
        self.lets = []

    def get(self, tok, depth=0):
        tok = tok.strip()
        if tok.startswith("%"):
            if tok in self.multi:
                return tok + "*"
            if tok in self.defs:
                return self.defs[tok]
            return tok
        return imm(tok)


def main():
    path, name = sys.argv[1], sys.argv[2]
    for a in sys.argv[3:]:
        if "=" in a:
            k, v = a.split("=")
            PARAM_NAMES[k] = v
    text = open(path).read()
    body = parse_kernel(text, name)
    lines = [l.strip() for l in body.splitlines()]
    sym = Sym()
    # first pass: count definitions
    defcount = {}
    inst_re = re.compile(r"^(@!?%p\d+\s+)?([a-z0-9_.]+)\s+(.*);$")
    for l in lines:
        m = inst_re.match(l)
        if not m:
            continue
        op, args = m.group(2), m.group(3)
        if op.startswith(("st.", "bra", "bar", "ret", "setp", "red", "atom")):
            if op.startswith("setp"):
                d = args.split(",")[0].strip()
                defcount[d] = defcount.get(d, 0) + 1
            continue
        dst = args.split(",")[0].strip()
        if dst.startswith("{"):
            for d in re.findall(r"%\w+", args.split("}")[0]):
                defcount[d] = defcount.get(d, 0) + 1
        else:
            defcount[dst] = defcount.get(dst, 0) + 1
    sym.multi = {r for r, c in defcount.items() if c > 1}

    out = []
    for l in lines:
        if l.endswith(":") and l.startswith("$L"):
            out.append("--- " + l)
            continue
        m = inst_re.match(l)
        if not m:
            continue
        pred, op, args = m.group(1) or "", m.group(2), m.group(3)
        # split args respecting braces / brackets
        parts = [a.strip() for a in re.split(r",\s*(?![^{]*\})", args)]
        g = sym.get
        if op.startswith("st.global") or op.startswith("st.shared") or op.startswith("st.local"):
            addr = parts[0]
            vals = parts[1]
            abase = re.findall(r"%\w+", addr)
            aexpr = addr
            if abase:
                aexpr = addr.replace(abase[0], g(abase[0]))
            if vals.startswith("{"):
                vs = [g(v) for v in re.findall(r"%\w+|0[fFdD][0-9A-Fa-f]+", vals)]
                out.append("%sSTORE %s %s <- %s" % (pred, op, aexpr, " | ".join(vs)))
            else:
                out.append("%sSTORE %s %s <- %s" % (pred, op, aexpr, g(vals)))
            continue
        if op.startswith(("red.", "atom.")):
            out.append("%sATOM %s %s" % (pred, op, ", ".join(g(p) if p.startswith("%") else p for p in parts)))
            continue
        if op.startswith("bra"):
            out.append("%sBRA %s" % (pred, args))
            continue
        if op.startswith("setp"):
            d = parts[0]
            cmp_ = op.split(".")[1]
            e = "(%s %s %s)" % (g(parts[1]), cmp_, g(parts[2]))
            if ".f32" in op or ".f64" in op:
                out.append("SETP %s = %s" % (d, e))
            sym.defs[d] = e
            continue
        dst = parts[0]
        srcs = parts[1:]
        e = None
        t = op.split(".")
        base = t[0]
        is_f = op.endswith(".f32") or op.endswith(".f64") or ".f32." in op or ".f64." in op
        suffix = "d" if "f64" in op else ""
        if op.startswith("ld.param"):
            e = "P" + srcs[0]
        elif op.startswith("ld."):
            addr = srcs[0]
            abase = re.findall(r"%\w+", addr)
            aexpr = addr
            if abase:
                aexpr = addr.replace(abase[0], g(abase[0]))
            space = t[1]
            if dst.startswith("{"):
                ds = re.findall(r"%\w+", dst)
                for i, d in enumerate(ds):
                    if d not in sym.multi:
                        sym.defs[d] = "LD.%s%s.%d" % (space, aexpr, i)
                continue
            e = "LD.%s%s" % (space, aexpr)
        elif base in ("mov", "cvta"):
            e = g(srcs[0])
        elif base == "cvt":
            e = "cvt<%s>(%s)" % (".".join(t[1:]), g(srcs[0]))
        elif base in ("add", "sub", "mul", "div", "min", "max") and is_f:
            sign = {"add": "+", "sub": "-", "mul": "*", "div": "/"}.get(base)
            if sign:
                e = "(%s %s%s %s)" % (g(srcs[0]), sign, suffix, g(srcs[1]))
            else:
                e = "%s%s(%s, %s)" % (base, suffix, g(srcs[0]), g(srcs[1]))
        elif base == "fma":
            e = "fma%s%s(%s, %s, %s)" % (suffix, "" if ".rn" in op else "<" + t[1] + ">", g(srcs[0]), g(srcs[1]), g(srcs[2]))
        elif base in ("neg", "abs", "sqrt", "rcp", "ex2", "lg2", "sin", "cos", "rsqrt") and is_f:
            e = "%s%s%s(%s)" % (base, suffix, "" if "approx" not in op else "~", g(srcs[0]))
        elif base == "selp":
            e = "sel(%s ? %s : %s)" % (g(srcs[2]), g(srcs[0]), g(srcs[1]))
        elif base in ("mul", "mad", "add", "sub", "shl", "shr", "and", "or", "xor", "not", "min", "max"):
            e = "%s(%s)" % (op, ", ".join(g(s) for s in srcs))
            if len(e) > 120:
                e = dst  # keep integer address soup short
        elif base == "cvt":
            e = "cvt(%s)" % g(srcs[0])
        else:
            e = "%s(%s)" % (op, ", ".join(g(s) for s in srcs))
            if len(e) > 160:
                e = dst
        if dst.startswith("{"):
            continue
        for k, v in PARAM_NAMES.items():
            e = e.replace("P[param%s]" % k, v)
        if dst in sym.multi:
            out.append("%sPHI %s := %s" % (pred, dst, e))
        elif len(e) > MAXLEN and dst.startswith("%f"):
            out.append("  %s = %s" % (dst, e))
            sym.defs[dst] = dst
        elif len(e) > MAXLEN:
            sym.defs[dst] = dst
        else:
            sym.defs[dst] = e
    print("\n".join(out))


if __name__ == "__main__":
    main()
