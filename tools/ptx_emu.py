#!/usr/bin/env python3
"""A small PTX interpreter for the integer subset our kernels use (development tool: there is no GPU in the build container).

Runs ONE thread (run) or one whole thread block with its barriers and shared memory (run_block) of a kernel from a `nvcc -ptx` file on
the CPU: add/sub with carry flags, mul/mad.wide, shifts, funnel shifts, brev, logic, setp/selp, predicated branches, ld/st.global,
ld/st.shared, bar.sync, mbarrier + cp.async.bulk (TMA staging, completed at once), ld.const, ld.param (scalars or a by-value struct given as bytes).  Enough to execute the Poseidon permutation, the field arithmetic and
the NTT butterflies exactly as nvcc emitted them (inline PTX included), so that device code can be compared with the oracle before any
GPU time is spent.  What it cannot see is ptxas (PTX -> SASS); the GPU parity tests remain the judge of that.

usage (library):  emu = PtxEmu(open("k.ptx").read()); emu.run("kernel_name_substring", params=[...], mem={addr: u64}, tid=0, ctaid=0, ntid=128)
                  emu.run_block("kernel", params, mem, ntid=32, ctaid=(bx, by))     # all threads, lock-step at bar.sync
"""
import re

M32, M64 = (1 << 32) - 1, (1 << 64) - 1
PARAM_BASE = 1 << 58          # fake addresses of kernel parameters whose address is taken (ld.param through a register)
LOCAL_BASE = 1 << 60          # per-thread stack (".local" depot): addresses at and above this go to a thread-private store


class PtxEmu:
    def __init__(self, text):
        self.consts = {}
        for m in re.finditer(r"\.const\s+\.align\s+\d+\s+\.b8\s+(\S+)\[(\d+)\]\s*=\s*\{([^}]*)\}", text):
            data = bytes(int(x) for x in m.group(3).split(","))
            self.consts[m.group(1)] = data + bytes(int(m.group(2)) - len(data))       # trailing zeros are not written out
        # constant arrays get fake base addresses so that pointer arithmetic on them works
        self.const_base = {n: (1 << 56) + (i << 32) for i, n in enumerate(self.consts)}
        # shared arrays (extern or static): offsets in one per-block shared window; extern arrays start at 0
        self.shared_base, off = {}, 0
        for m in re.finditer(r"\.shared\s+\.align\s+(\d+)\s+\.b8\s+(\w+)\[(\d*)\]", text):
            self.shared_base[m.group(2)] = 0 if not m.group(3) else off
            if m.group(3):
                off += (int(m.group(3)) + 15) // 16 * 16
        self.kernels, self.funcs = {}, {}
        text = re.sub(r"\.func[^{;]*;", "", text)          # forward declarations of device functions
        for m in re.finditer(r"\.(entry|func)\s+(?:\([^)]*\)\s*)?(\S+?)\s*\(\s*(.*?)\)\s*(?:\.\w+[^\{]*)?\{(.*?)\n\}", text, re.S):
            name, params, body = m.group(2), m.group(3), m.group(4)
            pnames = [re.sub(r"\[\d+\]$", "", p.strip().split()[-1]) for p in params.split(",") if p.strip()]
            (self.kernels if m.group(1) == "entry" else self.funcs)[name] = (pnames, self._parse(body))

    @staticmethod
    def _parse(body):
        body = re.sub(r"//[^\n]*", "", body)
        ins, labels = [], {}
        for raw in body.replace("\n", " ").split(";"):
            line = raw.strip()
            while True:
                m = re.match(r"(\$?\w+):(?!:)\s*(.*)", line)      # compiler labels ($L__BB0_1:) and the plain ones of inline asm
                if not m:
                    break
                labels[m.group(1)] = len(ins)
                line = m.group(2).strip()
            if not line or line.startswith(".") or line in ("{", "}"):
                # declarations (.reg ...) carry no semantics here
                m = re.match(r"[{}\s]*(.*)", line)
                line = m.group(1) if m else ""
                if not line or line.startswith("."):
                    continue
            line = line.strip("{} \t")
            if not line or line.startswith("."):          # "{ .param .b64 param0" opening a call sequence
                continue
            pred = None
            m = re.match(r"@(!?)(%?\w+)\s+(.*)", line)             # @%p3, @!%p3, and @p of an inline-asm predicate
            if m:
                pred, line = (m.group(2), m.group(1) == "!"), m.group(3)
            parts = line.split(None, 1)
            op = parts[0]
            args = []
            if op.startswith("call"):         # call.uni (retval), fname, (param0, ...)   |   call.uni fname, (param0, ...)
                m = re.match(r"(?:\(\s*(\w+)\s*\)\s*,\s*)?(\S+?)\s*(?:,\s*\((.*)\))?\s*$", parts[1])
                ins.append((pred, "call", [m.group(1), m.group(2), [x.strip() for x in (m.group(3) or "").split(",") if x.strip()]]))
                continue
            if len(parts) > 1:
                depth, cur = 0, ""
                for ch in parts[1]:
                    if ch == "{":
                        depth += 1
                    if ch == "}":
                        depth -= 1
                    if ch == "," and depth == 0:
                        args.append(cur.strip()); cur = ""
                    else:
                        cur += ch
                if cur.strip():
                    args.append(cur.strip())
            ins.append((pred, op, args))
        return ins, labels

    # ---- execution ---------------------------------------------------------------------------------------------
    def run(self, kernel, params, mem, tid=0, ctaid=0, ntid=128, max_steps=50_000_000, nctaid=1):
        self.nctaid = nctaid
        """one thread; barriers are ignored (single-thread kernels or kernels whose barriers only order other threads' data)"""
        for _ in self._exec(kernel, params, mem, {}, tid, ctaid, ntid, max_steps):
            pass
        return mem

    def run_block(self, kernel, params, mem, ntid, ctaid=0, max_steps=50_000_000):
        """all `ntid` threads of one block: every thread runs to its next bar.sync (or to its end), then the next phase starts"""
        smem = {}
        threads = [self._exec(kernel, params, mem, smem, t, ctaid, ntid, max_steps) for t in range(ntid)]
        live = list(range(ntid))
        while live:
            nxt = []
            for t in live:
                try:
                    next(threads[t])
                    nxt.append(t)
                except StopIteration:
                    pass
            assert not nxt or len(nxt) == len(live), "threads disagree on a barrier"
            live = nxt
        return mem

    def _exec(self, kernel, params, mem, smem, tid, ctaid, ntid, max_steps):
        cx, cy = (ctaid, 0) if isinstance(ctaid, int) else ctaid
        name = [k for k in self.kernels if kernel in k]
        assert len(name) == 1, name
        pnames, (ins, labels) = self.kernels[name[0]]
        yield from self._body(ins, labels, dict(zip(pnames, params)), mem, smem, {}, tid, (cx, cy), ntid, [max_steps], 0)

    def _body(self, code, labels, pval, mem, smem, lmem, tid, cta, ntid, budget, depth, out=None):
        """one function activation: registers and call-sequence parameters are its own; `lmem` (the thread's stack) is shared with its
        callers, each call depth has its own depot window"""
        cx, cy = cta
        ins = code
        R, cc = {}, 0
        ptemp = {}                    # .param temporaries of call sequences and this function's return value, as bytearrays
        pbase = {n: PARAM_BASE + (i << 32) for i, n in enumerate(pval)}

        def param_bytes(b, off, n):
            if b.startswith("%"):     # the address of a parameter was taken (mov.b64 %rd, kernel_param_0)
                ad = val(b) + off
                b = [k for k, ba in pbase.items() if ba <= ad < ba + (1 << 32)][0]
                off = ad - pbase[b]
            pv = ptemp[b] if b in ptemp else pval[b]
            if isinstance(pv, (bytes, bytearray)):
                assert off + n <= len(pv), (b, off, n)
                return int.from_bytes(pv[off:off + n], "little")
            assert off == 0
            return pv & ((1 << (8 * n)) - 1)

        # memories are dicts of little-endian 8-byte words keyed by their (8-aligned) byte address; accesses are naturally aligned
        def rd(d, ad, nbytes):
            w = d[ad & ~7]
            return (w >> ((ad & 7) * 8)) & ((1 << (8 * nbytes)) - 1)

        def wr(d, ad, nbytes, v):
            sh, m = (ad & 7) * 8, (1 << (8 * nbytes)) - 1
            d[ad & ~7] = (d.get(ad & ~7, 0) & ~(m << sh) & M64) | ((v & m) << sh)

        def space(ad):
            return lmem if ad >= LOCAL_BASE else mem

        def val(a, bits=64):
            a = a.strip()
            if a.startswith("%"):
                if a == "%tid.x":
                    return tid
                if a == "%ctaid.x":
                    return cx
                if a == "%ctaid.y":
                    return cy
                if a == "%ntid.x":
                    return ntid
                if a == "%nctaid.x":
                    return getattr(self, "nctaid", 1)
                return R[a]
            if a in self.const_base:
                return self.const_base[a]
            if a in pbase:
                return pbase[a]
            if a in self.shared_base:
                return self.shared_base[a]
            if a.startswith("__local_depot"):
                return LOCAL_BASE + (depth << 32)
            if a.startswith("0x") or a.startswith("-0x"):
                return int(a, 16) & ((1 << bits) - 1)
            if a.endswith("U"):
                a = a[:-1]
            return int(a) & ((1 << bits) - 1)

        def addr(a):
            m = re.match(r"\[(.+?)(?:\+(-?\d+))?\]", a)
            base, off = m.group(1), int(m.group(2) or 0)
            return base, off

        pc = 0
        try:
            while pc < len(ins):
                budget[0] -= 1
                assert budget[0] > 0, "step limit"
                pred, op, a = ins[pc]
                pc += 1
                if pred is not None and bool(R[pred[0]]) == pred[1]:
                    continue
                o = op.split(".")
                base = o[0]
                if op == "ret":
                    break
                if base == "bar":
                    yield
                    continue
                if base == "mbarrier":
                    # TMA staging (ntt_pass_tma_kernel): the bulk copies below complete at once, so init / arrive.expect_tx carry no
                    # state here; try_wait is where every thread waits for ALL threads' copies — a phase boundary of the lock-step
                    # model, exactly like bar.sync — and then sees the phase complete
                    if o[1] == "try_wait":
                        yield
                        R[a[0]] = 1
                    continue
                if base == "fence":
                    continue
                if base == "cp" and o[1:3] == ["async", "bulk"]:      # cp.async.bulk.shared::cluster.global...bytes [smem], [gmem], size, [mbar]
                    db, doff = addr(a[0])
                    sb, soff = addr(a[1])
                    n = val(a[2])
                    assert n % 8 == 0 and (val(db) + doff) % 16 == 0 and (val(sb) + soff) % 16 == 0, "bulk copies are 16-byte aligned"
                    for i in range(0, n, 8):
                        wr(smem, (val(db) + doff + i) & M32, 8, rd(mem, (val(sb) + soff + i) & M64, 8))
                    continue
                if base == "bra":
                    pc = labels[a[0]]
                    continue
                if base == "call":
                    ret, fname, fargs = a
                    fp, (fins, flabels) = self.funcs[fname]
                    assert len(fp) == len(fargs), fname
                    callee = {}
                    yield from self._body(fins, flabels, {n: bytes(ptemp[x]) for n, x in zip(fp, fargs)}, mem, smem, lmem, tid, cta, ntid,
                                          budget, depth + 1, callee)
                    if ret:
                        ptemp[ret] = callee["func_retval0"]
                    continue
                mt = re.match(r"[usbf](\d+)$", o[-1])
                bits = int(mt.group(1)) if mt else 32
                mask = (1 << bits) - 1
                if base == "mov":
                    if a[0].startswith("{"):      # mov.b64 {lo, hi}, x
                        lo, hi = [x.strip() for x in a[0].strip("{}").split(",")]
                        v = val(a[1]); R[lo], R[hi] = v & M32, v >> 32
                    elif a[1].startswith("{"):    # mov.b64 x, {lo, hi}
                        lo, hi = [x.strip() for x in a[1].strip("{}").split(",")]
                        R[a[0]] = (val(lo) & M32) | ((val(hi) & M32) << 32)
                    elif o[-1] == "pred":
                        R[a[0]] = val(a[1])
                    else:
                        R[a[0]] = val(a[1], bits) & mask
                elif base == "ld":
                    b, off = addr(a[1])
                    if o[1] == "param":               # scalars, by-value structs (bytes), call-sequence temporaries
                        n = bits // 8
                        if a[0].startswith("{"):
                            for i, d in enumerate(x.strip() for x in a[0].strip("{}").split(",")):
                                R[d] = param_bytes(b, off + i * n, n)
                        else:
                            R[a[0]] = param_bytes(b, off, n)
                    elif o[1] == "shared":
                        R[a[0]] = rd(smem, (val(b) + off) & M32, bits // 8)
                    elif o[1] == "const":
                        if b in self.consts:
                            data = self.consts[b]
                        else:                     # register-relative: find the array the address falls into
                            ad = val(b) + off
                            nm = [n for n, ba in self.const_base.items() if ba <= ad < ba + len(self.consts[n])]
                            assert len(nm) == 1, hex(ad)
                            data, off = self.consts[nm[0]], ad - self.const_base[nm[0]]
                        assert 0 <= off and off + bits // 8 <= len(data)
                        R[a[0]] = int.from_bytes(data[off:off + bits // 8], "little")
                    elif a[0].startswith("{"):        # vector load: consecutive elements
                        for i, d in enumerate(x.strip() for x in a[0].strip("{}").split(",")):
                            ad = (val(b) + off + i * bits // 8) & M64
                            R[d] = rd(space(ad), ad, bits // 8)
                    else:
                        ad = (val(b) + off) & M64
                        v = rd(space(ad), ad, bits // 8)
                        if o[-1][0] == "s" and v >> (bits - 1):       # ld.s32 into a 64-bit register sign-extends
                            v = (v - (1 << bits)) & (M64 if a[0].startswith("%rd") else M32)
                        R[a[0]] = v
                elif base == "atom":              # atom.global.{add,or,...}.{u64,b32,...} d, [a], b   (threads run one after the other)
                    b, off = addr(a[1])
                    ad = (val(b) + off) & M64
                    old_v = rd(space(ad), ad, bits // 8) if (ad & ~7) in space(ad) else 0
                    y = val(a[2], bits)
                    new_v = {"add": old_v + y, "or": old_v | y, "and": old_v & y, "xor": old_v ^ y, "max": max(old_v, y), "min": min(old_v, y),
                             "exch": y}[o[2]]
                    wr(space(ad), ad, bits // 8, new_v & mask)
                    R[a[0]] = old_v
                elif base == "st" and o[1] == "param":
                    b, off = addr(a[0])
                    buf = ptemp.setdefault(b, bytearray())
                    n = bits // 8
                    if len(buf) < off + n:
                        buf.extend(bytes(off + n - len(buf)))
                    buf[off:off + n] = (val(a[1], bits) & mask).to_bytes(n, "little")
                elif base == "st" and o[1] == "shared":
                    b, off = addr(a[0])
                    wr(smem, (val(b) + off) & M32, bits // 8, val(a[1]))
                elif base == "st":
                    b, off = addr(a[0])
                    if a[1].startswith("{"):
                        for i, d in enumerate(x.strip() for x in a[1].strip("{}").split(",")):
                            ad = (val(b) + off + i * bits // 8) & M64
                            wr(space(ad), ad, bits // 8, val(d))
                    else:
                        ad = (val(b) + off) & M64
                        wr(space(ad), ad, bits // 8, val(a[1]))
                elif base == "cvta":
                    R[a[0]] = val(a[1])
                elif base == "cvt":
                    sb, db = int(o[-1][1:]), int(o[-2][1:])      # cvt.<dst>.<src>: read at the source width (sign-extended for s), keep the destination width
                    v = val(a[1], 64) & ((1 << sb) - 1)
                    if o[-1][0] == "s" and v >> (sb - 1):
                        v -= 1 << sb
                    R[a[0]] = v & ((1 << db) - 1)
                elif base in ("add", "sub", "addc", "subc"):
                    x, y = val(a[1], bits), val(a[2], bits)
                    cin = cc if base in ("addc", "subc") else 0
                    if base in ("add", "addc"):
                        r = x + y + cin
                        cout = r >> bits
                    else:
                        r = x - y - cin
                        cout = 1 if r < 0 else 0
                    if "cc" in o:
                        cc = cout
                    R[a[0]] = r & mask
                elif base in ("div", "rem"):
                    assert o[-1][0] == "u", op
                    x, y = val(a[1], bits), val(a[2], bits)
                    R[a[0]] = (x // y if base == "div" else x % y) if y else mask
                elif base == "brev":
                    R[a[0]] = int("{:032b}".format(val(a[1], 32))[::-1], 2)
                elif base == "neg":
                    R[a[0]] = (-val(a[1], bits)) & mask
                elif base == "mul":
                    if o[1] == "wide":
                        R[a[0]] = (val(a[1], 32) * val(a[2], 32)) & M64
                    elif o[1] == "lo":
                        R[a[0]] = (val(a[1], bits) * val(a[2], bits)) & mask
                    elif o[1] == "hi" and o[-1][0] == "u":
                        R[a[0]] = (val(a[1], bits) * val(a[2], bits)) >> bits
                    else:
                        raise NotImplementedError(op)
                elif base == "mad":
                    if o[1] == "lo":
                        R[a[0]] = (val(a[1], bits) * val(a[2], bits) + val(a[3], bits)) & mask
                    else:
                        assert o[1] == "wide" and o[-1] == "u32", op
                        R[a[0]] = (val(a[1], 32) * val(a[2], 32) + val(a[3], 64)) & M64
                elif base == "shl":
                    s = val(a[2], 32)
                    R[a[0]] = (val(a[1], bits) << s) & mask if s < bits else 0
                elif base == "shr":
                    s = val(a[2], 32)
                    assert o[-1] in ("u32", "u64", "b32", "b64"), op
                    R[a[0]] = (val(a[1], bits) >> s) if s < bits else 0
                elif base == "shf":           # shf.{l,r}.wrap.b32 d, lo, hi, c
                    lo, hi, c = val(a[1], 32), val(a[2], 32), val(a[3], 32) & 31
                    v = (hi << 32) | lo
                    R[a[0]] = ((v << c) >> 32) & M32 if o[1] == "l" else (v >> c) & M32
                elif base in ("and", "or", "xor"):
                    x, y = val(a[1], bits), val(a[2], bits)
                    R[a[0]] = (x & y) if base == "and" else (x | y) if base == "or" else (x ^ y)
                elif base == "not":
                    R[a[0]] = (0 if val(a[1]) else 1) if o[-1] == "pred" else (~val(a[1], bits)) & mask
                elif base == "setp":
                    x, y = val(a[1], bits), val(a[2], bits)
                    if o[-1].startswith("s"):
                        sx = x - (1 << bits) if x >> (bits - 1) else x
                        sy = y - (1 << bits) if y >> (bits - 1) else y
                        x, y = sx, sy
                    R[a[0]] = int({"gt": x > y, "ge": x >= y, "lt": x < y, "le": x <= y, "eq": x == y, "ne": x != y}[o[1]])
                elif base in ("min", "max"):
                    x, y = val(a[1], bits), val(a[2], bits)
                    if o[-1][0] == "s":
                        x, y = (x - (1 << bits) if x >> (bits - 1) else x), (y - (1 << bits) if y >> (bits - 1) else y)
                    R[a[0]] = (min(x, y) if base == "min" else max(x, y)) & mask
                elif base == "selp":
                    R[a[0]] = val(a[1], bits) if R[a[3]] else val(a[2], bits)
                else:
                    raise NotImplementedError(op)
        except (KeyError, AssertionError, NotImplementedError, IndexError) as e:
            if not getattr(e, "ptx_where", None):     # innermost activation only
                e.ptx_where = True
                e.args = e.args + ("depth %d, instruction %d: %s %s" % (depth, pc - 1, ins[pc - 1][1], ins[pc - 1][2]),
                                   {r: hex(R[r]) for x in ins[pc - 1][2] if isinstance(x, str) for r in re.findall(r"%\w+", x) if r in R})
            raise
        if out is not None:           # a device function: hand the return value (st.param [func_retval0]) to the caller
            out.update(ptemp)
