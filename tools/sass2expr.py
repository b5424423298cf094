#!/usr/bin/env python3
"""sass2expr -- symbolic read-out of the floating-point dataflow of one SASS kernel.

Development tool (not part of the product or the oracle).  ptxas fuses `mul.f32`+`add.f32`
pairs that nvcc left un-contracted in PTX, so the arithmetic a kernel really performs is only
visible in SASS.  This tool walks `cuobjdump -sass` output in program order, keeps a
register -> expression map, and prints an expression for every store / compare, so that the
reference kernels' exact evaluation order (which products are fused into FFMAs) can be
reproduced with explicit intrinsics in our kernels and with fmaf() in the CPU oracle.

It recognises the inline IEEE division / sqrt / reciprocal expansions (MUFU.RCP / MUFU.RSQ +
FFMA fix-up + slow-path call) and folds them back into div() / sqrt() / rcp().

usage: cuobjdump -sass -fun <mangled> obj.o | sass2expr.py [name=c[0x0][0x390] ...]
"""
import re
import sys

MAXLEN = 90
names = {}
for a in sys.argv[1:]:
    if "=" in a:
        k, v = a.split("=", 1)
        names[v] = k

inst_re = re.compile(r"^\s*/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)\s*(.*?)\s*;")

regs = {}      # "R12" -> expr
load_alias = {}  # address string -> short name
tmp_id = [0]
out = []
pending_div = None   # (pred, a, b)
last_ffma_dst = [None]
skip_until = [None]


def fmt_const(tok):
    tok = tok.strip()
    if tok in names:
        return names[tok]
    return tok


def val(tok):
    tok = tok.strip()
    neg = False
    absv = False
    tok = tok.replace(".reuse", "")
    if tok.startswith("-"):
        neg = True
        tok = tok[1:]
    if tok.startswith("|") and tok.endswith("|"):
        absv = True
        tok = tok[1:-1]
    if tok.startswith("-"):
        neg = not neg
        tok = tok[1:]
    m = re.match(r"^(U?R\d+)(\.64|\.H[01]|\.B[0-3])?$", tok)
    if tok == "RZ" or tok == "URZ":
        e = "0"
    elif m:
        r = m.group(1)
        e = regs.get(r, r)
    else:
        e = fmt_const(tok)
    if absv:
        e = "abs(%s)" % e
    if neg:
        e = "-%s" % e if re.match(r"^[\w.\[\]]+$", e) else "-(%s)" % e
    return e


def setreg(dst, e, pred=""):
    dst = dst.strip().replace(".reuse", "")
    m = re.match(r"^(U?R\d+)", dst)
    if not m:
        return
    r = m.group(1)
    if pred:
        old = regs.get(r, r)
        e = "sel(%s ? %s : %s)" % (pred.strip(), e, old)
    if len(e) > MAXLEN:
        tmp_id[0] += 1
        t = "t%d" % tmp_id[0]
        out.append("  %s = %s" % (t, e))
        e = t
    regs[r] = e


def split_ops(s):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch == "[":
            depth += 1
        elif ch == "]":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        parts.append(cur.strip())
    return parts


for line in sys.stdin:
    m = inst_re.match(line)
    if not m:
        continue
    addr, pred, op, rest = m.group(1), (m.group(2) or "").strip(), m.group(3), m.group(4)
    if skip_until[0] is not None:
        if int(addr, 16) < skip_until[0]:
            continue
        skip_until[0] = None
    ops = split_ops(rest)
    base = op.split(".")[0]
    if base in ("FMUL", "FADD", "FFMA", "DMUL", "DADD", "DFMA"):
        d = "d" if base[0] == "D" else ""
        if base[1:] == "MUL":
            e = "(%s *%s %s)" % (val(ops[1]), d, val(ops[2]))
            if val(ops[1]) == "0" or val(ops[2]) == "0":
                e = "0"
        elif base[1:] == "ADD":
            e = "(%s +%s %s)" % (val(ops[1]), d, val(ops[2]))
            if val(ops[2]) == "0":
                e = val(ops[1])
            elif val(ops[1]) == "0":
                e = val(ops[2])
        else:
            x, y, z = val(ops[1]), val(ops[2]), val(ops[3])
            if x == "0" or y == "0":
                e = z                      # fma(0, y, z) == z for finite y
            elif z == "0":
                e = "(%s *%s %s)" % (x, d, y)
            else:
                e = "fma%s(%s, %s, %s)" % (d, x, y, z)
            if "rcp~" in e or any("rcp~" in str(regs.get(re.sub(r"[^UR0-9]", "", o.replace(".reuse", "")), ""))
                                  for o in ops[1:4]):
                last_ffma_dst[0] = ops[0]
        if ".FTZ" in op or ".RM" in op or ".RP" in op or ".SAT" in op:
            e = op.split(".", 1)[1] + ":" + e
        setreg(ops[0], e, pred)
    elif base == "MUFU":
        setreg(ops[0], "%s~(%s)" % (op.split(".")[1].lower(), val(ops[1])), pred)
    elif base == "FCHK":
        pending_div = (ops[0], val(ops[1]), val(ops[2]))
    elif base == "BRA" and pred and pending_div and pred.lstrip("@!") == pending_div[0]:
        # end of the inline division fast path: the last FFMA result is a / b (correctly rounded)
        if last_ffma_dst[0]:
            setreg(last_ffma_dst[0], "div(%s, %s)" % (pending_div[1], pending_div[2]))
        pending_div = None
        # skip the out-of-line slow path (MOVs + CALL + MOV result) up to the branch target
        mt = re.search(r"0x([0-9a-f]+)", rest)
        if mt:
            skip_until[0] = int(mt.group(1), 16)
    elif base in ("FSETP", "DSETP"):
        out.append("SETP %s %s = %s(%s, %s)  [%s]" % (pred, ops[0], op.split(".", 1)[1], val(ops[2]), val(ops[3]),
                                                      ",".join(ops[4:])))
    elif base == "FMNMX":
        e = "%s(%s, %s)" % ("min" if ops[-1] in ("PT",) else ("max" if ops[-1] == "!PT" else "mnmx[" + ops[-1] + "]"),
                            val(ops[1]), val(ops[2]))
        setreg(ops[0], e, pred)
    elif base in ("F2F", "I2F", "F2I", "I2FP", "F2FP", "FRND"):
        setreg(ops[0], "cvt<%s>(%s)" % (op.split(".", 1)[1] if "." in op else op, val(ops[1])), pred)
    elif base == "FSEL":
        setreg(ops[0], "fsel(%s ? %s : %s)" % (ops[3], val(ops[1]), val(ops[2])), pred)
    elif base in ("MOV", "UMOV"):
        setreg(ops[0], val(ops[1]), pred)
    elif base in ("IMAD", "UIMAD") and ".MOV" in op:
        setreg(ops[0], val(ops[-1]), pred)
    elif base in ("LDG", "LD", "LDS", "LDL"):
        a = "%s%s" % (base, re.sub(r"(U?R\d+)", lambda mm: "{" + str(regs.get(mm.group(1), mm.group(1)))[:60] + "}",
                                   ops[1]))
        if a not in load_alias:
            load_alias[a] = "L%d" % (len(load_alias) + 1)
            out.append("  %s := %s" % (load_alias[a], a))
        d0 = ops[0].strip()
        mreg = re.match(r"^R(\d+)", d0)
        width = 2 if ".64" in op else (4 if ".128" in op else 1)
        if width == 1 or not mreg:
            setreg(ops[0], load_alias[a], pred)
        else:
            for i in range(width):
                setreg("R%d" % (int(mreg.group(1)) + i), "%s.%d" % (load_alias[a], i), pred)
    elif base in ("LDC", "LDCU", "ULDC"):
        setreg(ops[0], fmt_const(ops[1]), pred)
    elif base in ("STG", "ST", "STS", "STL"):
        a = re.sub(r"(U?R\d+)", lambda mm: "{" + str(regs.get(mm.group(1), mm.group(1)))[:60] + "}", ops[0])
        out.append("%s STORE %s %s <- %s" % (pred, op, a, val(ops[1])))
    elif base in ("BRA", "EXIT", "CALL", "RET", "BSYNC", "BSSY"):
        if base in ("BRA", "EXIT", "CALL"):
            out.append("%s %s %s" % (pred, op, rest))
    else:
        # integer / misc: keep a short opaque expression so address arithmetic stays readable
        if ops:
            e = "%s(%s)" % (op, ", ".join(val(o) for o in ops[1:]))
            if len(e) > 70:
                e = "%s@%s" % (base, addr)
            setreg(ops[0], e, pred)

print("\n".join(out))
