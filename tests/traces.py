"""Synthetic trace generators (numpy) for the STARK tables: valid traces for the tables whose row semantics are simple
enough to restate here, random traces for throughput.  Layouts: (ncols, n) uint64, column-major.

Sources: LogicStark rows /root/reference/evm_arithmetization/src/logic.rs:165-237; MemoryContinuationStark
memory_continuation_stark.rs:54-100; MemoryStark columns memory/columns.rs:10-51 + generate_first_change_flags_and_rc
memory_stark.rs:134-200; padding CPU rows generation/mod.rs:646-663."""
import numpy as np

P = 0xFFFFFFFF00000001
T_ARITHMETIC, T_BYTE_PACKING, T_CPU, T_KECCAK, T_KECCAK_SPONGE, T_LOGIC, T_MEMORY, T_MEM_BEFORE, T_MEM_AFTER = range(9)
NUM_COLUMNS = {T_ARITHMETIC: 116, T_BYTE_PACKING: 71, T_CPU: 85, T_KECCAK: 2431, T_KECCAK_SPONGE: 438, T_LOGIC: 523,
               T_MEMORY: 30, T_MEM_BEFORE: 12, T_MEM_AFTER: 12}


def random_trace(table, log_n, seed):
    rng = np.random.default_rng(seed)
    return rng.integers(0, P, size=(NUM_COLUMNS[table], 1 << log_n), dtype=np.uint64)


def memcont_trace(log_n, seed, fill=0.7):
    """filter in {0,1}; address and value limbs arbitrary 32-bit values"""
    rng = np.random.default_rng(seed)
    n = 1 << log_n
    t = rng.integers(0, 1 << 32, size=(12, n), dtype=np.uint64)
    t[0] = (rng.random(n) < fill).astype(np.uint64)
    return t


def logic_trace(log_n, seed, fill=0.8):
    """random AND/OR/XOR rows, the rest padding (all-zero rows)"""
    rng = np.random.default_rng(seed)
    n = 1 << log_n
    t = np.zeros((523, n), dtype=np.uint64)
    nops = int(n * fill)
    op = rng.integers(0, 3, size=nops)
    a = rng.integers(0, 2, size=(256, nops), dtype=np.uint64)
    b = rng.integers(0, 2, size=(256, nops), dtype=np.uint64)
    for k in range(3):
        t[k, :nops] = (op == k)
    t[3:259, :nops] = a
    t[259:515, :nops] = b
    res = np.where(op == 0, a & b, np.where(op == 1, a | b, a ^ b))
    w = (np.uint64(1) << np.arange(32, dtype=np.uint64))[:, None]
    for l in range(8):
        t[515 + l, :nops] = (res[32 * l:32 * l + 32] * w).sum(axis=0)
    return t


def logic_ops(nops, seed):
    """random operations in the layout zkgpu_logic_generate_trace takes: (nops, 9) uint64 = operator, input0 limbs, input1 limbs"""
    rng = np.random.default_rng(seed)
    ops = rng.integers(0, 1 << 64, size=(nops, 9), dtype=np.uint64)
    ops[:, 0] = rng.integers(0, 3, size=nops).astype(np.uint64)
    return ops


def logic_trace_from_ops(log_n, ops):
    """LogicStark::generate_trace (logic.rs:165-237) from explicit operations (same layout as logic_ops)"""
    ops = np.ascontiguousarray(ops, dtype=np.uint64).reshape(-1, 9)
    n, nops = 1 << log_n, ops.shape[0]
    assert nops <= n
    t = np.zeros((523, n), dtype=np.uint64)
    for k in range(3):
        t[k, :nops] = (ops[:, 0] == k)
    for l in range(4):
        a, b = ops[:, 1 + l], ops[:, 5 + l]
        t[3 + 64 * l:3 + 64 * l + 64, :nops] = _bits(a)
        t[259 + 64 * l:259 + 64 * l + 64, :nops] = _bits(b)
        r = np.where(ops[:, 0] == 0, a & b, np.where(ops[:, 0] == 1, a | b, a ^ b))
        t[515 + 2 * l, :nops] = r & np.uint64(0xFFFFFFFF)
        t[515 + 2 * l + 1, :nops] = r >> np.uint64(32)
    return t


def memory_trace_simple(log_n):
    """A valid MemoryStark trace made of dummy reads at (0, 0, r), timestamp 0: every ordering, initialisation and
    range-check constraint holds with range_check == 0 and all frequencies on counter 0."""
    n = 1 << log_n
    t = np.zeros((30, n), dtype=np.uint64)
    t[3] = 1                                   # is_read
    t[6] = np.arange(n, dtype=np.uint64)       # addr_virtual
    t[17] = 1                                  # virtual_first_change (also on the last row: wraps to row 0)
    t[20] = 34 * 35                            # preinitialized_segments_aux = (0-34)(0-35)
    t[28] = np.arange(n, dtype=np.uint64)      # counter
    t[29, 0] = n                               # frequencies: range_check == 0 on every row
    return t


# ---- CpuStark: all-padding trace (generation/mod.rs:646-663; halt.rs:16-52) ----------------------------------------------
def cpu_padding_trace(log_n, halt_final=0x1234):
    n = 1 << log_n
    t = np.zeros((85, n), dtype=np.uint64)
    t[2] = halt_final                                  # program_counter
    t[3] = 7                                           # stack_len (constant)
    t[4] = 1                                           # is_kernel_mode
    t[5] = 1000                                        # gas (constant)
    t[40] = np.arange(1, n + 1, dtype=np.uint64)       # clock starts at 1
    return t


def cpu_program_trace(log_n, program, halt_final=0x1234, gas0=50, seed=0, inputs=(), log=None, keccak=None, load32=None, store32=None,
                      syscall_jumptable=0x4000, exception_jumptable=0x5000, jumptable=None, stale=None, push32=None):
    """CpuStark trace with ACTIVE rows: a kernel-mode program that runs into `halt_final`, then the padding rows.
    program: string of  J JUMPDEST 0x5b | P PC 0x58 | 0 PUSH0 0x5f | N NOT 0x19 | X POP 0x50 | Z ISZERO 0x15 | E EQ 0x14 | A ADD 0x01 |
    M MUL 0x02 | S SUB 0x03 | D DIV 0x04 | O MOD 0x06 | L LT 0x10 | G GT 0x11 | B BYTE 0x1a | & AND 0x16 | "|" OR 0x17 | ^ XOR 0x18 |
    a ADDMOD 0x08 | m MULMOD 0x09 | u v w q DUP1 DUP2 DUP3 DUP16 | s t y SWAP1 SWAP2 SWAP16 | j JUMP 0x56 | i JUMPI 0x57 |
    f g h ADDFP254 MULFP254 SUBFP254 0x0c-0x0e | K KECCAK_GENERAL 0x21 | I PROVER_INPUT 0xee (the next word of `inputs`,
    else a random word) | C GET_CONTEXT 0xf6 | < SHL 0x1b | > SHR 0x1c (displacement on top) | l MLOAD_GENERAL 0xfb | r MSTORE_GENERAL 0xfc (address word = virtual | segment << 32 | context << 64) |
    R MLOAD_32BYTES 0xf8 (address word on top, length below it; pushes the bytes packed big-endian: load32(address word, length, clock) -> bytes,
    else random bytes) | W V T MSTORE_32BYTES_32 / _5 / _1 0xdf 0xc4 0xc0 (address word on top, value < 256^n below it; pushes the address
    word + n; store32(address word, value, n, clock) is told about the write) |
    e EXIT_KERNEL 0xf9 (pops kexit_info = pc | kernel flag << 32 | gas << 192; this model stays in kernel mode) |
    Y a SYSCALL row (opcode 0x20) and x an EXCEPTION row (exc_stop, code 6 — the one exception a kernel-mode row may raise; faulting opcode
    0xfe): both read their handler address from the jump table (3 bytes at syscall_jumptable + 3 opcode / exception_jumptable + 3 code,
    through BytePacking: jumptable(virt, handler, clock) is told), push kexit_info and continue at the handler = the first JUMPDEST at
    least two instructions further on, in kernel mode with gas 0. |
    c SET_CONTEXT 0xf7 (contextops.rs:150-215, operation.rs:371-454): pops new context << 64 | pruning flag; every context has its own stack,
    whose length is kept in its ContextMetadata::StackSize cell while another context runs; `stale` (a list) receives the pruned contexts. |
    p.. PUSH2 0x61 with its two immediate bytes (the two characters after it are never executed; their values are drawn from the seed).
    USER MODE: an EXIT_KERNEL whose kexit_info has kernel flag 0 leaves kernel mode; there p, X, N, J and Y may run (the code is read from
    (context, Segment::Code), a user-mode PUSH reads its immediate through BytePacking — push32(code context, address, bytes, clock) is
    told —, pushes are bounds-checked against MAX_USER_STACK_SIZE) until a syscall row Y brings the kernel back.
    The string is the CODE (instruction c at address halt_final - len + c); execution starts at its first instruction and follows the
    jumps until it reaches halt_final (jump targets are built on the stack from PC values, e.g. "PPS" pushes 1).
    log: a list that receives (instruction, operands..., result) of every arithmetic / logic instruction executed.
    The stack starts empty; the model keeps the 256-bit words so that the cached top (mem_channels[0]), the second-operand reads
    (mem_channels[1]), the partial-channel write of the old top and the new-top read after POP carry consistent values.
    Exercises decode.rs, control_flow.rs, gas.rs, clock.rs, stack.rs (every StackBehavior shape: push, no-op, unary, binary, pop with and
    without a new-top read, stack_inv / stack_inv_aux / stack_inv_aux_2), pc.rs, push0.rs, simple_logic/{not,eq_iszero}.rs, membus.rs,
    halt.rs with operation flags set.  (The results of the arithmetic and logic instructions are checked by the Arithmetic / Logic
    tables through cross-table lookups only; in this table they meet the decode, gas and stack constraints.)"""
    n = 1 << log_n
    k = len(program)
    assert 0 < k < n
    M256 = (1 << 256) - 1
    t = np.zeros((85, n), dtype=np.uint64)
    t[4] = 1                                           # is_kernel_mode (the executed rows set their own below)
    t[40] = np.arange(1, n + 1, dtype=np.uint64)       # clock
    opcode = {"J": 0x5b, "P": 0x58, "0": 0x5f, "N": 0x19, "X": 0x50, "Z": 0x15, "E": 0x14, "A": 0x01, "M": 0x02,
              "S": 0x03, "D": 0x04, "O": 0x06, "L": 0x10, "G": 0x11, "B": 0x1a, "&": 0x16, "|": 0x17, "^": 0x18, "a": 0x08, "m": 0x09,
              "j": 0x56, "i": 0x57}
    flag = {"J": 13, "P": 21, "0": 21, "N": 11, "X": 11, "Z": 9, "E": 9, "A": 6, "M": 6,
            "S": 6, "D": 6, "O": 6, "L": 6, "G": 6, "B": 6, "&": 10, "|": 10, "^": 10, "a": 7, "m": 7, "j": 14, "i": 14}
    cost = {"J": 1, "P": 2, "0": 2, "N": 3, "X": 2, "Z": 3, "E": 3, "A": 3, "M": 5,
            "S": 3, "D": 5, "O": 5, "L": 3, "G": 3, "B": 3, "&": 3, "|": 3, "^": 3, "a": 8, "m": 8, "j": 8, "i": 10}
    dup, swap = {"u": 0, "v": 1, "w": 2, "q": 15}, {"s": 0, "t": 1, "y": 15}          # DUP1 / DUP2 / DUP3 / DUP16, SWAP1 / SWAP2 / SWAP16
    for c, i in dup.items():
        opcode[c], flag[c], cost[c] = 0x80 + i, 16, 3
    for c, i in swap.items():
        opcode[c], flag[c], cost[c] = 0x90 + i, 16, 3
    for c, (oc, fl) in {"f": (0x0c, 8), "g": (0x0d, 8), "h": (0x0e, 8), "K": (0x21, 13), "I": (0xee, 15), "l": (0xfb, 20), "r": (0xfc, 20)}.items():   # kernel-only: no gas
        opcode[c], flag[c], cost[c] = oc, fl, 0
    opcode["C"], flag["C"], cost["C"] = 0xf6, 17, 0    # GET_CONTEXT (kernel-only)
    store_len = {"W": 32, "V": 5, "T": 1}              # MSTORE_32BYTES_n: opcode 0xc0 + n - 1 (decode.rs:206-211), kernel-only: no gas
    opcode["p"], flag["p"], cost["p"] = 0x61, 15, 3    # PUSH2 (gas.rs:121-125: G_VERYLOW)
    opcode["c"], flag["c"], cost["c"] = 0xf7, 17, 0    # SET_CONTEXT
    opcode["e"], flag["e"], cost["e"] = 0xf9, 19, 0    # EXIT_KERNEL
    opcode["Y"], flag["Y"], cost["Y"] = 0x20, 22, 0    # a syscall opcode (decode.rs does not constrain which: the handler checks)
    opcode["x"], flag["x"], cost["x"] = 0xfe, 23, 0    # exception raised at an invalid opcode
    opcode["R"], flag["R"], cost["R"] = 0xf8, 18, 0    # MLOAD_32BYTES
    for c, ln in store_len.items():
        opcode[c], flag[c], cost[c] = 0xc0 + ln - 1, 18, 0
    opcode["<"], flag["<"], cost["<"] = 0x1b, 12, 3    # SHL
    opcode[">"], flag[">"], cost[">"] = 0x1c, 12, 3    # SHR
    binary = {"<": lambda a, b: (b << a) & M256 if a < 256 else 0, ">": lambda a, b: b >> a if a < 256 else 0,
              "f": lambda a, b: (a + b) % BN_BASE, "g": lambda a, b: (a * b) % BN_BASE, "h": lambda a, b: (a - b) % BN_BASE,
              "K": lambda a, b: (a * 0x9E3779B97F4A7C15 + b) & M256,         # KECCAK_GENERAL: the digest comes from the sponge table (CTL)
              "A": lambda a, b: (a + b) & M256, "M": lambda a, b: (a * b) & M256, "S": lambda a, b: (a - b) & M256,
              "D": lambda a, b: a // b if b else 0, "O": lambda a, b: a % b if b else 0, "L": lambda a, b: int(a < b), "G": lambda a, b: int(a > b),
              "B": lambda a, b: (b >> (8 * (31 - a))) & 0xFF if a < 32 else 0, "&": lambda a, b: a & b, "|": lambda a, b: a | b, "^": lambda a, b: a ^ b}
    limbs = lambda x: [(x >> (32 * i)) & 0xFFFFFFFF for i in range(8)]
    inputs = list(inputs)                              # words PROVER_INPUT supplies, in order (random words once exhausted)
    base = halt_final - k                              # the program occupies the addresses base .. halt_final - 1
    pc, gas, ctx, kernel = base, gas0, 0, 1
    stacks, saved_sp = {0: []}, {}                     # per context: the stack, and the StackSize cell written when the context was left
    stack = stacks[0]
    read_top_next = False                              # the previous instruction was a POP / JUMP / JUMPI that left a non-empty stack
    r = -1
    while pc != halt_final:
        r += 1
        assert base <= pc < halt_final and r < n - 1, "the program left its code or does not halt"
        ins = program[pc - base]
        next_pc = pc + 1
        sl = len(stack)
        t[0, r], t[1, r], t[2, r], t[3, r], t[4, r], t[5, r] = ctx, (1 - kernel) * ctx, pc, sl, kernel, gas
        assert kernel or ins in "pXNJY", "instruction %r is not modelled in user mode" % ins
        next_kernel = kernel
        if stack:
            t[46:54, r] = limbs(stack[-1])             # mem_channels[0].value: the cached top of the stack
        if read_top_next:                              # ... read from memory at the start of this row (stack.rs:371-386)
            t[41, r], t[42, r], t[43, r], t[44, r], t[45, r] = 1, 1, ctx, 1, sl - 1
            read_top_next = False
        for b in range(8):
            t[24 + b, r] = (opcode[ins] >> b) & 1      # opcode_bits, little endian
        t[flag[ins], r] = 1
        if ins == "p":                                 # PUSH2: is_not_kernel, the immediate, the bounds check of a user-mode push
            imm = bytes(int(x) for x in np.random.default_rng(seed + 3000 + r).integers(0, 256, size=2))
            t[32, r] = 1 - kernel                      # general.push().is_not_kernel (control_flow.rs:88-90)
            if sl > 0:
                t[80, r], t[81, r], t[82, r], t[83, r], t[84, r] = 1, 0, ctx, 1, sl - 1
                t[36, r], t[37, r] = pow(sl, P - 2, P), 1
            if not kernel:
                t[39, r] = pow((sl + 1 - 1025) % P, P - 2, P)                               # stack_len_bounds_aux (stack.rs:328-333)
                if push32 is not None:
                    push32((1 - kernel) * ctx, pc + 1, imm, r + 1)
            stack.append(int.from_bytes(imm, "big"))
            next_pc = pc + 3
        elif ins in "P0":                              # push only
            if sl > 0:                                 # the old top goes to memory through the partial channel
                t[80, r], t[81, r], t[82, r], t[83, r], t[84, r] = 1, 0, ctx, 1, sl - 1
                t[36, r], t[37, r] = pow(sl, P - 2, P), 1                                   # general.stack(): stack_inv, stack_inv_aux
            stack.append(pc if ins == "P" else 0)
        elif ins in "NX":                              # not_pop: stack_inv / stack_inv_aux refer to stack_len - 1
            assert sl >= 1
            aux = int(sl != 1)
            t[36, r], t[37, r] = (pow(sl - 1, P - 2, P) if aux else 0), aux
            if ins == "N":
                stack[-1] ^= M256
            else:
                t[38, r] = aux                         # stack_inv_aux_2 = stack_inv_aux * (1 - opcode_bits[0])
                stack.pop()
                read_top_next = bool(aux)
        elif ins == "Z":
            assert sl >= 1
            x = limbs(stack[-1])
            nz = [l for l in x if l]
            for i, l in enumerate(x):                  # general.logic().diff_pinv (eq_iszero.rs:25-42)
                t[32 + i, r] = pow(l, P - 2, P) * pow(len(nz), P - 2, P) % P if l else 0
            stack[-1] = int(stack[-1] == 0)
        elif ins in "ji":                              # jumps.rs:67-175 JUMP(dst) = JUMPI(dst, 1); kernel mode: the JUMPDEST bit is not read
            npop = 1 if ins == "j" else 2
            assert sl >= npop
            dst = stack.pop()
            if ins == "j":
                cond = 1
                t[59, r] = 1                           # mem_channels[1].value = 1 with the channel unused
            else:
                cond = stack.pop()
                t[54, r], t[55, r], t[56, r], t[57, r], t[58, r] = 1, 1, ctx, 1, sl - 2
                t[59:67, r] = limbs(cond)
            csum = sum(limbs(cond)) % P
            t[32, r], t[33, r] = int(cond != 0), (pow(csum, P - 2, P) if csum else 0)          # general.jumps(): should_jump, cond_sum_pinv
            aux = int(sl != npop)
            t[36, r], t[37, r] = (pow(sl - npop, P - 2, P) if aux else 0), aux                 # general.stack(): stack_inv, stack_inv_aux
            t[67, r], t[68, r], t[69, r], t[70, r], t[71, r] = 0, 1, ctx, 14, dst & 0xFFFFFFFF   # JUMPDEST-bit channel: unused in kernel mode
            t[72, r] = 1
            read_top_next = bool(aux)
            if cond:
                assert dst < (1 << 32)
                next_pc = dst
        elif ins in dup:                               # dup_swap.rs:112-141: the top goes to memory (channel 1), element n is read (channel 2)
            i = dup[ins]
            assert sl > i
            t[54, r], t[55, r], t[56, r], t[57, r], t[58, r] = 1, 0, ctx, 1, sl - 1
            t[59:67, r] = limbs(stack[-1])
            t[67, r], t[68, r], t[69, r], t[70, r], t[71, r] = 1, 1, ctx, 1, sl - 1 - i
            t[72:80, r] = limbs(stack[-1 - i])
            stack.append(stack[-1 - i])
        elif ins in swap:                              # dup_swap.rs:207-244: element n+1 is read (channel 1), the top is written there (channel 2)
            i = swap[ins]
            assert sl > i + 1
            t[54, r], t[55, r], t[56, r], t[57, r], t[58, r] = 1, 1, ctx, 1, sl - 2 - i
            t[59:67, r] = limbs(stack[-2 - i])
            t[67, r], t[68, r], t[69, r], t[70, r], t[71, r] = 1, 0, ctx, 1, sl - 2 - i
            t[72:80, r] = limbs(stack[-1])
            stack[-1], stack[-2 - i] = stack[-2 - i], stack[-1]
        elif ins in "am":                              # three operands: the second and third are read through mem_channels[1], [2]
            assert sl >= 3
            a, b, c = stack.pop(), stack.pop(), stack.pop()
            t[54, r], t[55, r], t[56, r], t[57, r], t[58, r] = 1, 1, ctx, 1, sl - 2
            t[59:67, r] = limbs(b)
            t[67, r], t[68, r], t[69, r], t[70, r], t[71, r] = 1, 1, ctx, 1, sl - 3
            t[72:80, r] = limbs(c)
            stack.append(0 if c == 0 else ((a + b) % c if ins == "a" else (a * b) % c))
            if log is not None:
                log.append((ins, a, b, c, stack[-1]))
        elif ins == "c":                               # SET_CONTEXT
            assert sl >= 1
            w = limbs(stack.pop())
            assert w[0] in (0, 1) and not any(w[i] for i in (1, 3, 4, 5, 6, 7))
            new_ctx = w[2]
            saved_sp[ctx] = sl - 1                     # old stack pointer -> (ctx, ContextMetadata, StackSize), a CTL-only memory write
            new_sp = saved_sp.get(new_ctx, 0)          # new stack pointer <- (new_ctx, ContextMetadata, StackSize): 0 for a fresh context
            new_stack = stacks.setdefault(new_ctx, [])
            assert len(new_stack) == new_sp
            t[32, r] = w[0]                            # general.context_pruning().pruning_flag
            if w[0] and stale is not None:
                stale.append(ctx)
            aux = int(new_sp != 0)
            t[36, r], t[37, r], t[38, r] = (pow(new_sp, P - 2, P) if aux else 0), aux, aux
            if aux:                                    # the new top is read through channel 2 and handed to the next row's channel 0
                t[67, r], t[68, r], t[69, r], t[70, r], t[71, r] = 1, 1, new_ctx, 1, new_sp - 1
                t[72:80, r] = limbs(new_stack[-1])
            ctx, stack = new_ctx, new_stack
        elif ins == "e":                               # EXIT_KERNEL (jumps.rs:13-65): pc, kernel flag and gas come from the popped kexit_info
            assert sl >= 1
            info = limbs(stack.pop())
            assert info[1] in (0, 1) and info[7] == 0
            aux = int(sl != 1)
            t[36, r], t[37, r] = (pow(sl - 1, P - 2, P) if aux else 0), aux
            read_top_next = bool(aux)
            next_pc, gas, next_kernel = info[0], info[6], info[1]
            if not next_kernel:                        # exit_kernel is in MIGHT_OVERFLOW (stack.rs:23-44): the stack it returns with is bounds-checked
                t[39, r] = pow((sl - 1 - 1025) % P, P - 2, P)
        elif ins in "Yx":                              # syscalls_exceptions.rs:23-134 (operation.rs:735-810, 983-1080)
            code = 6 if ins == "x" else None           # exc_stop
            virt = (exception_jumptable + 3 * code) if ins == "x" else (syscall_jumptable + 3 * opcode[ins])
            handler = next(base + i for i in range(pc - base + 2, k) if program[i] == "J")
            if ins == "x":
                for b in range(3):
                    t[32 + b, r] = (code >> b) & 1     # general.exception().exc_code_bits
            t[54, r], t[55, r], t[56, r], t[57, r], t[58, r] = 0, 1, 0, 0, virt      # the jump-table channel: described, not used (BytePacking reads it)
            t[59, r] = handler
            old_top = stack[-1] if stack else 0
            if sl > 0:                                 # push: the old top goes to memory through the partial channel
                t[80, r], t[81, r], t[82, r], t[83, r], t[84, r] = 1, 0, ctx, 1, sl - 1
                t[36, r], t[37, r] = pow(sl, P - 2, P), 1
            info = (pc + 1 if ins == "Y" else pc) | (kernel << 32) | (gas << 192)
            stack.append(info)
            if jumptable is not None:
                jumptable(virt, handler, r + 1)
            if log is not None:
                log.append((ins, opcode[ins], old_top, handler, info))
            next_pc, gas, next_kernel = handler, 0, 1
        elif ins == "C":                               # GET_CONTEXT (contextops.rs:82-102, 277-301): pushes context << 64; the old top goes out through channel 2
            if sl > 0:
                t[67, r], t[68, r], t[69, r], t[70, r], t[71, r] = 1, 0, ctx, 1, sl - 1
                t[72:80, r] = limbs(stack[-1])
                t[36, r], t[37, r] = pow(sl, P - 2, P), 1
            stack.append(ctx << 64)
        elif ins == "l":                               # MLOAD_GENERAL (memio.rs:22-57): the address word on top, the loaded word replaces it
            assert sl >= 1
            virt, seg, actx = limbs(stack[-1])[:3]     # get_addr (cpu_stark.rs:318-323)
            val = int.from_bytes(np.random.default_rng(seed + 1000 + r).bytes(32), "little")
            t[54, r], t[55, r], t[56, r], t[57, r], t[58, r] = 1, 1, actx, seg, virt
            t[59:67, r] = limbs(val)
            aux = int(sl != 2)                         # the stack view of every m_op_general row refers to stack_len - 2 (memio.rs:170-176)
            t[36, r], t[37, r] = (pow((sl - 2) % P, P - 2, P) if aux else 0), aux
            stack[-1] = val
        elif ins == "r":                               # MSTORE_GENERAL (memio.rs:137-200): value on top, address word below it
            assert sl >= 2
            stack.pop()
            addr = stack.pop()
            virt, seg, actx = limbs(addr)[:3]
            t[54, r], t[55, r], t[56, r], t[57, r], t[58, r] = 1, 1, ctx, 1, sl - 2
            t[59:67, r] = limbs(addr)
            t[80, r], t[81, r], t[82, r], t[83, r], t[84, r] = 1, 0, actx, seg, virt      # the store goes through the partial channel
            aux = int(sl != 2)
            t[36, r], t[37, r], t[38, r] = (pow(sl - 2, P - 2, P) if aux else 0), aux, aux
            read_top_next = bool(aux)
        elif ins == "I":                               # PROVER_INPUT: pushes whatever the prover supplies (push behaviour; is_not_kernel = 0)
            if sl > 0:
                t[80, r], t[81, r], t[82, r], t[83, r], t[84, r] = 1, 0, ctx, 1, sl - 1
                t[36, r], t[37, r] = pow(sl, P - 2, P), 1
            stack.append(inputs.pop(0) if inputs else int.from_bytes(np.random.default_rng(seed + r).bytes(32), "little"))
            if log is not None:
                log.append(("I", stack[-2] if len(stack) > 1 else 0, stack[-1]))
        elif ins in "EAMSDOLGB&|^fghK<>RWVT":          # two operands: the second one is read through mem_channels[1]
            assert sl >= 2
            a, b = stack.pop(), stack.pop()
            t[54, r], t[55, r], t[56, r], t[57, r], t[58, r] = 1, 1, ctx, 1, sl - 2
            t[59:67, r] = limbs(b)
            if ins == "E":
                d = [(x - y) % P for x, y in zip(limbs(a), limbs(b))]
                nz = [x for x in d if x]
                for i, x in enumerate(d):
                    t[32 + i, r] = pow(x, P - 2, P) * pow(len(nz), P - 2, P) % P if x else 0
                stack.append(int(a == b))
            elif ins == "K" and keccak is not None:    # the digest of the bytes at address word a, length b (callback: the sponge operation)
                stack.append(keccak(a, b, r + 1))
            elif ins == "R":                           # MLOAD_32BYTES (operation.rs:893-931): b bytes at address word a, packed big-endian
                assert 1 <= b <= 32
                data = load32(a, b, r + 1) if load32 is not None else np.random.default_rng(seed + 2000 + r).bytes(b)
                assert len(data) == b
                stack.append(int.from_bytes(data, "big"))
            elif ins in store_len:                     # MSTORE_32BYTES_n (operation.rs:963-981, byte_unpacking.rs:11-44): pushes the advanced address word
                ln = store_len[ins]
                assert b < (1 << (8 * ln)) and (a >> 96) == 0 and (a & 0xFFFFFFFF) + ln < (1 << 32)
                if store32 is not None:
                    store32(a, b, ln, r + 1)
                stack.append(a + ln)
            else:
                if log is not None:
                    log.append((ins, a, b, binary[ins](a, b)))
                if ins in "<>":                        # shift.rs:14-60: 2^d is read from the kernel's shift table unless d >= 2^32
                    hi = sum(limbs(a)[1:]) % P
                    t[32, r] = pow(hi, P - 2, P) if hi else 0                                   # general.shift().high_limb_sum_inv
                    t[67, r], t[68, r], t[69, r], t[70, r], t[71, r] = int(hi == 0), int(hi == 0), 0, 13, a & 0xFFFFFFFF
                    if hi == 0:
                        t[72:80, r] = limbs((1 << a) if a < 256 else 0)
                if ins in "fgh":                       # modfp254.rs: the BN254 prime sits where the modulus of the general operations goes
                    t[72:80, r] = limbs(BN_BASE)
                stack.append(binary[ins](a, b))
        gas += cost[ins]
        pc, kernel = next_pc, next_kernel
    k = r + 1                                          # executed rows
    t[0, k:], t[2, k:], t[3, k:], t[5, k:] = ctx, halt_final, len(stack), gas
    if stack:
        for l, v in enumerate(limbs(stack[-1])):
            t[46 + l, k:] = v
    # a POP that leaves a non-empty stack reads the new top on the NEXT row, and a halt row may not use a memory channel (halt.rs:33-41)
    assert not read_top_next, "the program must not end with a POP that leaves a non-empty stack"
    return t


def cpu_demo_program():
    """a longer mix of every instruction cpu_program_trace knows -> (program, inputs)"""
    A = lambda c, sg, v: v | (sg << 32) | (c << 64)
    program = ("0PPPSuAuAPAiNJ"            # JUMPI taken over the N, [0] stays
               "X" "IIIIrlXJ" "X"          # MSTORE_GENERAL, MLOAD_GENERAL
               "II<I>" "C" "Z" "&"         # SHL, SHR, GET_CONTEXT, ISZERO, AND
               "IIfIgIh" "IK"              # FP254 operations, KECCAK_GENERAL
               "PPvwstuE|^" "IIIam"        # DUP / SWAP, EQ, OR, XOR, ADDMOD, MULMOD
               "IDIOILIGIBISIM" "NXJ")     # DIV MOD LT GT BYTE SUB MUL, NOT, POP
    inputs = [A(9, 9, 9), 1, A(2, 5, 77), 123, 0xABCDEF, 5, 3]
    return program, inputs


# ---- KeccakStark (keccak_stark.rs:70-250) ---------------------------------------------------------------------------------
KECCAK_RC = [0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000, 0x000000000000808B,
             0x0000000080000001, 0x8000000080008081, 0x8000000000008009, 0x000000000000008A, 0x0000000000000088,
             0x0000000080008009, 0x000000008000000A, 0x000000008000808B, 0x800000000000008B, 0x8000000000008089,
             0x8000000000008003, 0x8000000000008002, 0x8000000000000080, 0x000000000000800A, 0x800000008000000A,
             0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008]
KECCAK_R = [[0, 36, 3, 41, 18], [1, 44, 10, 45, 2], [62, 6, 43, 15, 61], [28, 55, 25, 21, 56], [27, 20, 39, 8, 14]]
K_TIMESTAMP, K_START_A = 24, 25
K_START_C = K_START_A + 50
K_START_C_PRIME = K_START_C + 320
K_START_A_PRIME = K_START_C_PRIME + 320
K_START_A_PP = K_START_A_PRIME + 1600
K_START_A_PP_BITS = K_START_A_PP + 50
K_A_PPP_00_LO = K_START_A_PP_BITS + 64


def _rotl(x, r):
    r %= 64
    if r == 0:
        return x
    return ((x << np.uint64(r)) | (x >> np.uint64(64 - r)))


def _bits(x):
    """(64, nperm) array of the bits of a uint64 vector"""
    return ((x[None, :] >> np.arange(64, dtype=np.uint64)[:, None]) & np.uint64(1)).astype(np.uint64)


def keccak_trace(log_n, inputs, timestamps=None):
    """inputs: (nperm, 25) uint64, lane i = y*5 + x.  Returns (trace (2431, n), outputs (nperm, 25))."""
    inputs = np.ascontiguousarray(inputs, dtype=np.uint64)
    nperm = inputs.shape[0]
    n = 1 << log_n
    assert nperm * 24 <= n
    t = np.zeros((2431, n), dtype=np.uint64)
    if timestamps is None:
        timestamps = np.arange(1, nperm + 1, dtype=np.uint64) * np.uint64(7)
    A = [[inputs[:, y * 5 + x].copy() for y in range(5)] for x in range(5)]
    M32 = np.uint64(0xFFFFFFFF)
    S32 = np.uint64(32)
    for rnd in range(24):
        rows = np.arange(nperm) * 24 + rnd
        t[rnd, rows] = 1
        t[K_TIMESTAMP, rows] = timestamps
        for x in range(5):
            for y in range(5):
                t[K_START_A + (x * 5 + y) * 2, rows] = A[x][y] & M32
                t[K_START_A + (x * 5 + y) * 2 + 1, rows] = A[x][y] >> S32
        C = [A[x][0] ^ A[x][1] ^ A[x][2] ^ A[x][3] ^ A[x][4] for x in range(5)]
        Cp = [C[x] ^ C[(x + 4) % 5] ^ _rotl(C[(x + 1) % 5], 1) for x in range(5)]
        for x in range(5):
            t[K_START_C + 64 * x:K_START_C + 64 * x + 64, rows] = _bits(C[x])
            t[K_START_C_PRIME + 64 * x:K_START_C_PRIME + 64 * x + 64, rows] = _bits(Cp[x])
        Ap = [[A[x][y] ^ C[x] ^ Cp[x] for y in range(5)] for x in range(5)]
        for x in range(5):
            for y in range(5):
                o = K_START_A_PRIME + x * 320 + y * 64
                t[o:o + 64, rows] = _bits(Ap[x][y])
        B = [[_rotl(Ap[(x + 3 * y) % 5][x], KECCAK_R[(x + 3 * y) % 5][x]) for y in range(5)] for x in range(5)]
        App = [[B[x][y] ^ (~B[(x + 1) % 5][y] & B[(x + 2) % 5][y]) for y in range(5)] for x in range(5)]
        for x in range(5):
            for y in range(5):
                t[K_START_A_PP + x * 10 + y * 2, rows] = App[x][y] & M32
                t[K_START_A_PP + x * 10 + y * 2 + 1, rows] = App[x][y] >> S32
        t[K_START_A_PP_BITS:K_START_A_PP_BITS + 64, rows] = _bits(App[0][0])
        a000 = App[0][0] ^ np.uint64(KECCAK_RC[rnd])
        t[K_A_PPP_00_LO, rows] = a000 & M32
        t[K_A_PPP_00_LO + 1, rows] = a000 >> S32
        App[0][0] = a000
        A = App
    out = np.stack([A[i % 5][i // 5] for i in range(25)], axis=1)
    return t, out


# ---- BytePackingStark (byte_packing_stark.rs:193-280) ------------------------------------------------------------------------
def byte_packing_trace(log_n, seed, fill=0.6):
    rng = np.random.default_rng(seed)
    n = 1 << log_n
    assert n >= 256
    t = np.zeros((71, n), dtype=np.uint64)
    nops = max(1, int(n * fill))
    lens = rng.integers(1, 33, size=nops)
    t[0, :nops] = rng.integers(0, 2, size=nops)                 # is_read
    t[1 + (lens - 1), np.arange(nops)] = 1                      # index_len[len-1]
    t[33, :nops] = rng.integers(0, 100, size=nops)              # addr_context
    t[34, :nops] = rng.integers(0, 30, size=nops)               # addr_segment
    t[35, :nops] = rng.integers(0, 1 << 20, size=nops)          # addr_virtual
    t[36, :nops] = rng.integers(1, 1 << 20, size=nops)          # timestamp
    byts = rng.integers(0, 256, size=(32, nops), dtype=np.uint64)
    byts[np.arange(32)[:, None] >= lens[None, :]] = 0           # bytes past the length are zero
    t[37:69, :nops] = byts
    t[69] = np.minimum(np.arange(n), 255).astype(np.uint64)     # range_counter
    t[70, :256] = np.bincount(t[37:69].astype(np.int64).ravel(), minlength=256).astype(np.uint64)
    return t


# ---- ArithmeticStark: ADD / SUB / LT / GT rows (addcy.rs:24-60) + range-check columns (arithmetic_stark.rs:130-156) -------------
def _limbs16(vals):
    """list of python ints (< 2^256) -> (16, len) uint64 of 16-bit limbs"""
    return np.array([[(v >> (16 * i)) & 0xFFFF for v in vals] for i in range(16)], dtype=np.uint64)


def arithmetic_addcy_trace(log_n, seed, nops=1000):
    rng = np.random.default_rng(seed)
    n = 1 << log_n
    assert n >= 1 << 16
    t = np.zeros((116, n), dtype=np.uint64)
    ops = rng.integers(0, 4, size=nops)           # 0 add, 1 sub, 2 lt, 3 gt
    a = [int.from_bytes(rng.bytes(32), "little") for _ in range(nops)]
    b = [int.from_bytes(rng.bytes(32), "little") for _ in range(nops)]
    M = 1 << 256
    out, aux = [], []
    for k in range(nops):
        if ops[k] == 0:
            s = a[k] + b[k]; out.append(s % M); aux.append(s >> 256)                 # in0 + in1 = out + cy 2^256
        elif ops[k] == 1:
            d = a[k] - b[k]; out.append(d % M); aux.append(1 if d < 0 else 0)        # in1 + out = in0 + cy 2^256
        elif ops[k] == 2:
            d = a[k] - b[k]; aux.append(d % M); out.append(1 if d < 0 else 0)        # in1 + aux = in0 + out 2^256
        else:
            d = b[k] - a[k]; aux.append(d % M); out.append(1 if d < 0 else 0)        # in0 + aux = in1 + out 2^256
    flag_col = {0: 0, 1: 2, 2: 11, 3: 12}
    for k in range(nops):
        t[flag_col[int(ops[k])], k] = 1
    t[18:34, :nops] = _limbs16(a)
    t[34:50, :nops] = _limbs16(b)
    t[66:82, :nops] = _limbs16(out)
    t[82:98, :nops] = _limbs16(aux)
    t[114] = np.minimum(np.arange(n), 65535).astype(np.uint64)
    t[115, :65536] = np.bincount(t[18:114].astype(np.int64).ravel(), minlength=65536).astype(np.uint64)
    return t


# ---- KeccakSpongeStark (keccak_sponge_stark.rs:253-543) -------------------------------------------------------------------------------
def _keccakf_u32s(state):
    """keccakf_u32s (cpu/kernel/keccak_util.rs:7-19): 50 u32 = 25 little-endian (lo, hi) lanes, lane i = x + 5y"""
    lanes = np.array([[int(state[2 * i]) | (int(state[2 * i + 1]) << 32) for i in range(25)]], dtype=np.uint64)
    out = keccak_trace(5, lanes)[1][0]
    res = []
    for i in range(25):
        res += [int(out[i]) & 0xFFFFFFFF, int(out[i]) >> 32]
    return res


def keccak_sponge_trace(log_n, ops):
    """KeccakSpongeStark::generate_trace.  ops: list of (context, segment, virt, timestamp, input bytes).
    Returns (trace (438, n), digests: the 32-byte Keccak-256 of every input as the rows compute it)."""
    n = 1 << log_n
    assert n >= 256
    rows, digests = [], []
    for ctx, seg, virt, ts, data in ops:
        data = bytes(data)
        state = [0] * 50
        absorbed = 0
        nblocks = len(data) // 136 + 1
        for b in range(nblocks):
            row = [0] * 438
            final = b == nblocks - 1
            block = list(data[136 * b:136 * b + 136])
            if final:                                          # generate_final_row: pad10*1, is_padding_byte
                k = len(block)
                block += [0] * (136 - k)
                if k == 135:
                    block[135] = 0b10000001
                else:
                    block[k] = 1
                    block[135] = 0b10000000
                for i in range(k, 136):
                    row[6 + i] = 1
            else:
                row[0] = 1                                     # is_full_input_block
            row[1:6] = [ctx, seg, virt, ts, absorbed]
            row[142:176] = state[:34]                          # original_rate_u32s
            row[176:192] = state[34:]                          # original_capacity_u32s
            row[192:328] = block                               # block_bytes
            for i in range(34):
                state[i] ^= int.from_bytes(bytes(block[4 * i:4 * i + 4]), "little")
            row[328:362] = state[:34]                          # xored_rate_u32s
            state = _keccakf_u32s(state)
            row[362:404] = state[8:]                           # partial_updated_state_u32s
            for l in range(8):
                for i in range(4):
                    row[404 + 4 * l + i] = (state[l] >> (8 * i)) & 0xFF      # updated_digest_state_bytes
            rows.append(row)
            absorbed += 136
        digests.append(b"".join(int(x).to_bytes(4, "little") for x in state[:8]))
    assert len(rows) <= n
    t = np.zeros((438, n), dtype=np.uint64)
    if rows:
        t[:, :len(rows)] = np.array(rows, dtype=np.uint64).T
    t[436] = np.minimum(np.arange(n), 255).astype(np.uint64)                 # range_counter
    t[437, :256] = np.bincount(t[192:328].astype(np.int64).ravel(), minlength=256).astype(np.uint64)   # rc_frequencies
    return t, digests


# ---- MemoryStark from operations (memory_stark.rs:104-462) -------------------------------------------------------------------------
MEM_OP_COLS = (0, 1, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14)       # filter, timestamp, is_read, context, segment, virtual, 8 value limbs


def memory_sorted_ops(log_n, seed, nops=40, stale=()):
    """A consistent random memory history in the form the host hands to the finishing step: operations sorted by (context, segment,
    virtual, timestamp), the (0, 0, 0) dummy read of fill_gaps in front, pad_memory_ops padding behind -> (14, n) uint64.
    Reads return the last value written (zero before any write); contexts 0..2, a few segments incl. the preinitialised 0 and 12."""
    rng = np.random.default_rng(seed)
    n = 1 << log_n
    mem, ops = {}, []
    for k in range(nops):
        addr = (int(rng.integers(0, 3)), int(rng.choice([0, 1, 2, 12])), int(rng.integers(0, 6)))
        ts = 2 + k
        if rng.integers(0, 2) or addr not in mem:
            val = [int(x) for x in rng.integers(0, 1 << 32, size=8)] if rng.integers(0, 4) else [0] * 8
            mem[addr] = val
            ops.append((addr, ts, 0, 1, val))
        else:
            ops.append((addr, ts, 1, 1, mem[addr]))
    ops.sort(key=lambda o: (o[0], o[1]))
    if ops[0][0][2] != 0 or ops[0][0][:2] != (0, 0):
        ops.insert(0, ((0, 0, 0), 1, 1, 0, [0] * 8))
    last = ops[-1]
    while len(ops) < n:
        ops.append(((last[0][0], last[0][1], last[0][2] + 1), last[1] + 1, 1, 0, [0] * 8))
    assert len(ops) == n
    t = np.zeros((14, n), dtype=np.uint64)
    for i, (addr, ts, is_read, filt, val) in enumerate(ops):
        t[:6, i] = [filt, ts, is_read, addr[0], addr[1], addr[2]]
        t[6:, i] = val
    return t


def memory_finish_reference(ops, stale=()):
    """Restatement of MemoryOp::into_row (timestamp_inv), generate_first_change_flags_and_rc, insert_stale_contexts and
    generate_trace_col_major (memory_stark.rs:104-131, 134-213, 387-404, 240-294) on sorted / padded operations (14, n) -> (30, n)."""
    ops = np.asarray(ops, dtype=np.uint64)
    n = ops.shape[1]
    t = [[0] * n for _ in range(30)]
    for k, c in enumerate(MEM_OP_COLS):
        t[c] = [int(v) for v in ops[k]]
    for i in range(n):
        j = 0 if i == n - 1 else i + 1
        t[2][i] = pow(t[1][i], P - 2, P) if t[1][i] else 0
        cfc = t[4][i] != t[4][j]
        sfc = t[5][i] != t[5][j] and not cfc
        vfc = t[6][i] != t[6][j] and not sfc and not cfc
        t[15][i], t[16][i], t[17][i] = int(cfc), int(sfc), int(vfc)
        if i == n - 1:
            rc = 0
        elif cfc:
            rc = (t[4][j] - t[4][i] - 1) % P
        elif sfc:
            rc = (t[5][j] - t[5][i] - 1) % P
        elif vfc:
            rc = (t[6][j] - t[6][i] - 1) % P
        else:
            rc = (t[1][j] - t[1][i]) % P
        t[27][i] = rc
        assert rc < n
        t[20][i] = (t[5][j] - 34) * (t[5][j] - 35) % P
        t[19][i] = (t[5][j] - 0) * (t[5][j] - 12) * t[20][i] % P
        t[18][i] = t[19][i] * (t[15][i] + t[16][i] + t[17][i]) * t[3][j] % P
    for ctx in stale:
        t[21][ctx] = ctx + 1
        t[22][ctx] = 1
    t[28] = list(range(n))
    for i in range(n):
        t[29][t[27][i]] += 1
        if t[15][i] == 1 or t[16][i] == 1:
            if i < n - 1:
                t[29][t[6][i + 1]] += 1
            else:
                t[29][0] += 1
        ctx = t[4][i]
        if ctx + 1 == t[21][ctx]:
            t[24][i] = 1
            t[23][ctx] += 1
        elif t[0][i] == 1 and (t[15][i] == 1 or t[16][i] == 1 or t[17][i] == 1):
            t[25][i] = 1
            if any(t[7 + l][i] for l in range(8)) or t[5][i] in (0, 12, 34, 35):
                t[26][i] = 1
    return np.array(t, dtype=np.uint64)


def arithmetic_mul_rows(a, b):
    """mul.rs:70-120 generate_mul for 256-bit a, b -> (output limbs, aux_lo limbs, aux_hi limbs), 16 x 16-bit each"""
    al = [(a >> (16 * i)) & 0xFFFF for i in range(16)]
    bl = [(b >> (16 * i)) & 0xFFFF for i in range(16)]
    prod = [sum(al[i] * bl[k - i] for i in range(k + 1)) for k in range(16)]          # pol_mul_lo
    out, cy = [], 0
    for k in range(16):
        t = prod[k] + cy
        cy = t >> 16
        out.append(t & 0xFFFF)
    d = [prod[k] - out[k] for k in range(16)]                                         # pol_sub_assign
    q = [0] * 16                                                                      # pol_remove_root_2exp
    q[0] = -(d[0] >> 16)
    for k in range(1, 15):
        q[k] = (q[k - 1] - d[k]) >> 16
    q[15] = -cy
    q = [c + (1 << 20) for c in q]                                                    # + AUX_COEFF_ABS_MAX
    assert all(0 <= c < (1 << 32) for c in q)
    return out, [c & 0xFFFF for c in q], [c >> 16 for c in q]


def arithmetic_byte_row(idx, val):
    """BYTE (byte.rs:108-200 generate): index decomposition, the multiplexer tree over the value's limbs, the inverse of the high-limb sum -> the row"""
    M = (1 << 256) - 1
    row = [0] * 116
    row[13] = 1                                                                   # IS_BYTE
    il = [(idx >> (16 * i)) & 0xFFFF for i in range(16)]
    row[18:34], row[34:50] = il, [(val >> (16 * i)) & 0xFFFF for i in range(16)]
    for i in range(5):
        row[82 + i] = (idx >> i) & 1                                              # BYTE_IDX_DECOMP
    row[87] = il[0] >> 5                                                          # BYTE_IDX_DECOMP_HI
    hi_sum = (row[87] + sum(il[1:])) % P
    inv = pow(hi_sum, P - 2, P) if hi_sum else 1
    row[91:95] = [(inv >> (16 * i)) & 0xFFFF for i in range(4)]                   # BYTE_IDX_HI_LIMB_SUM_INV_0..3
    row[90] = int(hi_sum != 0)                                                    # BYTE_IDX_IS_LARGE
    lvl, src, dest = 3, 34, 98
    while True:
        ln = 1 << lvl
        src += (0 if (idx >> (lvl + 1)) & 1 else 1) * ln
        row[dest:dest + ln] = row[src:src + ln]
        if lvl == 0:
            break
        src, dest, lvl = dest, dest + ln, lvl - 1
    lo, hi = row[dest] & 0xFF, row[dest] >> 8
    row[88], row[89] = lo << 8, hi                                                # BYTE_LAST_LIMB_LO (stored * 256), _HI
    row[113] = lo if idx & 1 else hi                                              # tree[15]
    out = row[113] if idx < 32 else 0
    assert out == ((val >> (8 * (31 - idx))) & 0xFF if idx < 32 else 0)
    row[66:82] = [(out >> (16 * i)) & 0xFFFF for i in range(16)]
    return row


def arithmetic_mul_trace(log_n, seed, nops=200):
    """ArithmeticStark trace of MUL, SHL and BYTE operations (arithmetic_stark.rs:158-190 + mul.rs / shift.rs / byte.rs generate), incl. edge operands"""
    rng = np.random.default_rng(seed)
    n = 1 << log_n
    assert n >= 1 << 16
    t = np.zeros((116, n), dtype=np.uint64)
    M = (1 << 256) - 1
    vals = [(0, 0), (1, M), (M, M), (M, 2), (1 << 255, 2), (0xFFFF, 0xFFFF)] + \
           [(int.from_bytes(rng.bytes(32), "little"), int.from_bytes(rng.bytes(32), "little")) for _ in range(nops)]
    for k, (a, b) in enumerate(vals):
        out, lo, hi = arithmetic_mul_rows(a, b)
        assert sum(o << (16 * i) for i, o in enumerate(out)) == (a * b) & M
        t[1, k] = 1                                                                   # IS_MUL
        t[18:34, k] = [(a >> (16 * i)) & 0xFFFF for i in range(16)]
        t[34:50, k] = [(b >> (16 * i)) & 0xFFFF for i in range(16)]
        t[66:82, k] = out
        t[82:98, k] = lo
        t[98:114, k] = hi
    # BYTE (byte.rs:108-200): index decomposition, the multiplexer tree over the value's limbs, the inverse of the high-limb sum
    kb = len(vals) + 7
    byte_cases = [(0, M), (31, 0x1234), (32, M), (1 << 16, M), ((1 << 255) + 3, M), (5 + (7 << 5), M)] + \
                 [(int(rng.integers(0, 40)), int.from_bytes(rng.bytes(32), "little")) for _ in range(40)]
    for j, (idx, val) in enumerate(byte_cases):
        k = kb + j
        row = arithmetic_byte_row(idx, val)
        t[:, k] = row
    # SHL (shift.rs:41-76): (shift, input, 1 << shift or 0) in the three input registers, the MUL machinery on registers 1 and 2
    L = lambda x: [(x >> (16 * i)) & 0xFFFF for i in range(16)]
    k0 = len(vals)
    for j, (sh, x) in enumerate([(0, M), (1, M), (255, 3), (256, 5), (1 << 200, 7), (17, int.from_bytes(rng.bytes(32), "little")),
                                 (200, int.from_bytes(rng.bytes(32), "little"))]):
        disp = 0 if sh > 255 else 1 << sh
        out, lo, hi = arithmetic_mul_rows(x, disp)
        assert sum(o << (16 * i) for i, o in enumerate(out)) == ((x << sh) & M if sh < 256 else 0)
        k = k0 + j
        t[14, k] = 1                                                                  # IS_SHL
        t[18:34, k], t[34:50, k], t[50:66, k] = L(sh), L(x), L(disp)
        t[66:82, k], t[82:98, k], t[98:114, k] = out, lo, hi
    t[114] = np.minimum(np.arange(n), 65535).astype(np.uint64)
    t[115, :65536] = np.bincount(t[18:114].astype(np.int64).ravel(), minlength=65536).astype(np.uint64)
    return t


def _limbs(x, k=16):
    return [(x >> (16 * i)) & 0xFFFF for i in range(k)]


def _modular_op(pol, m, is_div, is_sub=False):
    """modular.rs:200-338 generate_modular_op: pol = the 31 input coefficients, m = the modulus (DIV: the denominator).
    -> (output, quotient limbs (32), second row)"""
    constr = list(pol) + [0]
    modulus, mod_limbs, mod_is_zero = m, _limbs(m), 0
    if m == 0:
        mod_is_zero = 1
        if is_div:
            modulus = 1 << 256                     # "modulus_limbs don't play a role below": the quotient is zero
        else:
            modulus = 1
            mod_limbs[0] = 1
    inp = sum(c << (16 * i) for i, c in enumerate(constr))                                   # columns_to_bigint
    out = inp % modulus
    quot = (inp - out) // modulus
    out_l = _limbs(out)
    quot_l = _limbs(quot, 32) if quot >= 0 else [-c for c in _limbs(-quot, 32)]              # bigint_to_columns: sign on every limb
    out_aux_red = _limbs((1 << 256) - modulus + out)
    for i in range(16):
        constr[i] -= out_l[i]
    prod = [sum(quot_l[i] * mod_limbs[k - i] for i in range(32) if 0 <= k - i < 16) for k in range(47)]   # pol_mul_wide2
    assert all(x == 0 for x in prod[32:])
    for i in range(32):
        constr[i] -= prod[i]
    q = [0] * 32                                                                              # pol_remove_root_2exp
    q[0] = -(constr[0] >> 16)
    for k in range(1, 31):
        q[k] = (q[k - 1] - constr[k]) >> 16
    q = [c + (1 << 20) for c in q]
    assert all(0 <= c < (1 << 32) for c in q)
    row2 = [0] * 116
    row2[18:34] = out_aux_red                                                                 # MODULAR_OUT_AUX_RED
    row2[34] = mod_is_zero                                                                    # MODULAR_MOD_IS_ZERO
    row2[35:66] = [c & 0xFFFF for c in q[:31]]                                                # MODULAR_AUX_INPUT_LO
    row2[66:97] = [c >> 16 for c in q[:31]]                                                   # MODULAR_AUX_INPUT_HI
    row2[97] = mod_is_zero if is_div else 0                                                   # MODULAR_DIV_DENOM_IS_ZERO
    if is_sub:                                                                                # :297-322: a negative quotient is stored offset, its sign after it
        assert all(c == 0 for c in quot_l[16:]) and all(abs(c) <= 0xFFFF for c in quot_l[:16])
        if quot < 0:
            quot_l = [c + 0xFFFF for c in quot_l[:16]] + [1] + [0] * 15
    return out, quot_l, row2


BN_BASE = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47        # extension_tower.rs:25-30


def arithmetic_modular_rows(op, a, b, m):
    """ADDMOD / MULMOD (modular.rs:343-383 generate) and DIV / MOD (divmod.rs:24-87; m unused) -> (row1, row2, result)"""
    al, bl = _limbs(a), _limbs(b)
    row1 = [0] * 116
    row1[18:34], row1[34:50] = al, bl
    flags = {"addmod": 5, "mulmod": 6, "addfp254": 7, "mulfp254": 8, "subfp254": 9, "submod": 10}
    if op in flags:
        if op.endswith("fp254"):
            m = BN_BASE
        if op.startswith("add"):
            pol = [al[i] + bl[i] for i in range(16)] + [0] * 15                              # pol_add
        elif op.startswith("sub"):
            pol = [al[i] - bl[i] for i in range(16)] + [0] * 15                              # pol_sub
        else:
            pol = [sum(al[i] * bl[k - i] for i in range(16) if 0 <= k - i < 16) for k in range(31)]   # pol_mul_wide
        out, quot_l, row2 = _modular_op(pol, m, False, is_sub=op.startswith("sub"))
        row1[flags[op]] = 1
        row1[50:66] = _limbs(m)
        row1[66:82] = _limbs(out)                                                             # MODULAR_OUTPUT
        row1[82:114] = quot_l                                                                 # MODULAR_QUO_INPUT
        return row1, row2, out
    if op == "shr":                                                                           # shift.rs:41-83: a = shift, b = input
        disp = 0 if a > 255 else 1 << a
        out, quot_l, row2 = _modular_op(bl + [0] * 15, disp, True)                            # divmod on registers 1 and 2, filter IS_SHR
        result = sum(c << (16 * i) for i, c in enumerate(quot_l[:16]))
        row1[15] = 1                                                                          # IS_SHR
        row1[50:66] = _limbs(disp)
        row1[66:82] = _limbs(result)
        row1[82:98] = _limbs(out)
        return row1, row2, result
    out, quot_l, row2 = _modular_op(al + [0] * 15, b, op == "div")                            # pol_extend(numerator), modulus = denominator
    assert all(x == 0 for x in quot_l[16:])
    result = sum(c << (16 * i) for i, c in enumerate(quot_l[:16])) if op == "div" else out
    row1[3 if op == "div" else 4] = 1                                                         # IS_DIV / IS_MOD
    row1[66:82] = _limbs(result)                                                              # OUTPUT_REGISTER
    row1[82:98] = _limbs(out) if op == "div" else quot_l[:16]                                 # AUX_INPUT_REGISTER_0: the other of (quotient, remainder)
    return row1, row2, result


def arithmetic_operation_rows(op, a, b, c=0):
    """Operation::to_rows (arithmetic/mod.rs:226-340) for one operation -> (list of one or two 116-column rows, result): the same row builders
    the traces above use, behind the reference's operation names"""
    M = 1 << 256
    if op in ("add", "sub", "lt", "gt"):                         # addcy.rs:71-95 generate
        if op == "add":
            out, aux = (a + b) % M, (a + b) >> 256
        elif op == "sub":
            out, aux = (a - b) % M, int(a < b)
        elif op == "lt":
            out, aux = int(a < b), (a - b) % M
        else:
            out, aux = int(a > b), (b - a) % M
        row = [0] * 116
        row[{"add": 0, "sub": 2, "lt": 11, "gt": 12}[op]] = 1
        row[18:34], row[34:50], row[66:82], row[82:98] = _limbs(a), _limbs(b), _limbs(out), _limbs(aux)
        return [row], out
    if op == "mul":
        out, lo, hi = arithmetic_mul_rows(a, b)
        row = [0] * 116
        row[1] = 1
        row[18:34], row[34:50], row[66:82], row[82:98], row[98:114] = _limbs(a), _limbs(b), out, lo, hi
        return [row], sum(o << (16 * i) for i, o in enumerate(out))
    if op == "byte":
        row = arithmetic_byte_row(a, b)
        return [row], sum(o << (16 * i) for i, o in enumerate(row[66:82]))
    row1, row2, res = arithmetic_modular_rows(op, a, b, c)
    return [row1, row2], res


def arithmetic_trace_from_operations(ops):
    """ArithmeticStark::generate_trace (arithmetic_stark.rs:158-190): the rows of the operations in order, padded to RANGE_MAX = 2^16 rows,
    then the range-check counter and frequency columns -> (116 x 2^16 trace, first row of every operation)"""
    t = np.zeros((116, 1 << 16), dtype=np.uint64)
    r, first = 0, []
    for op in ops:
        rows, _ = arithmetic_operation_rows(*op)
        first.append(r)
        for row in rows:
            t[:, r] = row
            r += 1
    t[114] = np.minimum(np.arange(1 << 16), 65535).astype(np.uint64)
    t[115, :65536] = np.bincount(t[18:114].astype(np.int64).ravel(), minlength=65536).astype(np.uint64)
    return t, first


def arithmetic_modular_trace(log_n, seed, nops=60):
    """ArithmeticStark trace of ADDMOD / MULMOD / DIV / MOD operations (two rows each), incl. modulus / denominator 0 and 1 and maximal operands"""
    rng = np.random.default_rng(seed)
    n = 1 << log_n
    assert n >= 1 << 16
    t = np.zeros((116, n), dtype=np.uint64)
    M = (1 << 256) - 1
    r = lambda: int.from_bytes(rng.bytes(32), "little")
    cases = [("addmod", M, M, M), ("mulmod", M, M, M), ("mulmod", M, M, M - 1), ("addmod", 5, 6, 0), ("mulmod", 5, 6, 0), ("mulmod", r(), r(), 1),
             ("addmod", 1, 2, 7), ("mulmod", 0, r(), r()), ("div", r(), 0, 0), ("mod", r(), 0, 0), ("div", M, 1, 0), ("mod", M, M, 0), ("div", 7, 9, 0),
             ("div", M, 3, 0), ("mod", 1 << 255, (1 << 128) + 1, 0), ("submod", 3, 5, 7), ("submod", 5, 3, 7), ("submod", 0, M, M - 1), ("submod", 9, 9, 0),
             ("addfp254", BN_BASE - 1, BN_BASE - 1, 0), ("mulfp254", BN_BASE - 1, BN_BASE - 2, 0), ("subfp254", 1, BN_BASE - 1, 0),
             ("submod", r(), r(), r()), ("submod", r() >> 100, r(), r() >> 3), ("shr", 0, M, 0), ("shr", 255, M, 0), ("shr", 256, M, 0), ("shr", 1 << 77, 5, 0),
             ("shr", 13, r(), 0)] + \
            [(("addmod", "mulmod", "div", "mod")[k % 4], r(), r() >> int(rng.integers(0, 200)) if k % 4 >= 2 else r(), r() >> int(rng.integers(0, 200)))
             for k in range(nops)]
    for k, (op, a, b, m) in enumerate(cases):
        row1, row2, out = arithmetic_modular_rows(op, a, b, m)
        if op == "shr":
            want = b >> a if a < 256 else 0
        elif op == "div":
            want = 0 if b == 0 else a // b
        elif op == "mod":
            want = 0 if b == 0 else a % b
        else:
            mm = BN_BASE if op.endswith("fp254") else m
            f = (lambda x, y: x + y) if op.startswith("add") else (lambda x, y: x - y) if op.startswith("sub") else (lambda x, y: x * y)
            want = 0 if mm == 0 else f(a, b) % mm
        assert out == want, (op, a, b, m)
        t[:, 2 * k] = row1
        t[:, 2 * k + 1] = row2
    t[114] = np.minimum(np.arange(n), 65535).astype(np.uint64)
    t[115, :65536] = np.bincount(t[18:114].astype(np.int64).ravel(), minlength=65536).astype(np.uint64)
    return t


# ---- a VALID multi-table segment: CPU padding + empty Arithmetic + Memory initialised from MemBefore, final state in MemAfter ----
def memory_trace_from_mem_before(log_n, addrs, values):
    """MemoryStark trace whose only real operations are the timestamp-0 initialisation writes of `mem_before`
    (memory_stark.rs:408-423): one row per (context, segment, virt) address, sorted, then dummy-read padding rows
    (pad_memory_ops, memory_stark.rs:358-385).  All addresses share context 0 and one non-preinitialised segment, with
    consecutive virtual addresses, so every range-checked difference is 0.  values: (K, 8) 32-bit limbs."""
    n = 1 << log_n
    K = len(addrs)
    assert 0 < K < n
    ctx, seg, v0 = addrs[0]
    assert all(a == (ctx, seg, v0 + i) for i, a in enumerate(addrs)) and seg not in (0, 12, 34, 35)
    t = np.zeros((30, n), dtype=np.uint64)
    p = lambda x: np.uint64(x % P)
    t[0, :K] = 1                                        # filter
    t[1, K:] = 1; t[2, K:] = 1                          # padding rows: timestamp 1, timestamp_inv 1
    t[3, K:] = 1                                        # is_read (initialisation rows are writes)
    t[4] = ctx; t[5] = seg
    t[6, :K] = v0 + np.arange(K); t[6, K:] = v0 + K
    t[7:15, :K] = np.asarray(values, dtype=np.uint64).T
    t[17, :K] = 1                                       # virtual_first_change: the next row has another address
    aux = (seg - 34) * (seg - 35)
    pre = seg * (seg - 12) * aux
    t[20] = p(aux); t[19] = p(pre)
    t[18, K - 1] = p(pre)                               # initialize_aux: the row after the last write is a first read
    t[25, :K] = 1; t[26, :K] = 1                        # maybe_in_mem_after, mem_after_filter
    t[28] = np.arange(n, dtype=np.uint64)               # counter
    t[29, 0] = n                                        # every range_check value is 0
    return t


def memcont_trace_from(log_n, addrs, values):
    n = 1 << log_n
    K = len(addrs)
    assert K <= n
    t = np.zeros((12, n), dtype=np.uint64)
    t[0, :K] = 1
    a = np.asarray(addrs, dtype=np.uint64)
    t[1, :K], t[2, :K], t[3, :K] = a[:, 0], a[:, 1], a[:, 2]
    t[4:12, :K] = np.asarray(values, dtype=np.uint64).T
    return t


def valid_segment(seed=0, log_cpu=6, log_mem=6, log_memcont=7, k=40, halt_final=0x1234):
    """traces[table] or None: Arithmetic (no operations), Cpu (padding), Memory, MemBefore, MemAfter in use; the optional
    BytePacking / Keccak / KeccakSponge / Logic tables left out (table_in_use = false)."""
    rng = np.random.default_rng(seed)
    addrs = [(0, 1, 100 + i) for i in range(k)]
    values = rng.integers(0, 1 << 32, size=(k, 8), dtype=np.uint64)
    tr = [None] * 9
    tr[T_ARITHMETIC] = arithmetic_addcy_trace(16, seed, nops=0)
    tr[T_CPU] = cpu_padding_trace(log_cpu, halt_final)
    tr[T_MEMORY] = memory_trace_from_mem_before(log_mem, addrs, values)
    tr[T_MEM_BEFORE] = memcont_trace_from(log_memcont, addrs, values)
    tr[T_MEM_AFTER] = memcont_trace_from(log_memcont, addrs, values)
    return tr


def cpu_segment(program, log_cpu=7, log_mem=9, log_memcont=7, log_logic=5, seed=0, k_before=6, num_channels=5, sponge_ops=None, packing_ops=None, inputs=(), keccak_inputs=None,
                extra_memory_rows=()):
    """A VALID multi-table segment around an executing Cpu program (no PROVER_INPUT, shifts, general memory or Keccak instructions: their
    lookups need more tables): Cpu (cpu_program_trace), Arithmetic (a row pair / row per MUL, DIV, MOD, ADDMOD, MULMOD executed), Logic (a row
    per AND / OR / XOR), Memory (every memory operation the Cpu rows send: the opcode fetch of every cycle, the general-purpose channels,
    the partial channel — plus the MemBefore initialisation writes), MemBefore, MemAfter.  Written from the reference's lookup definitions
    (cpu_stark.rs:324-379 mem_time_and_channel / ctl_data_code_memory / ctl_data_gp_memory / ctl_data_partial_memory with NUM_CHANNELS = 5,
    membus.rs:39; memory_stark.rs:35-95; all_stark.rs:153-417), so that a verifying segment checks OUR descriptors against them.
    `extra_memory_rows`: memory operations that no table sends — the kernel's writes of the public values (verifier.rs:536-737), each the 13
    values of the memory lookup; the Memory lookup then balances only with the verifier's extra looking sum (verifier.rs:319-512).
    -> (traces[9], labels)"""
    rng = np.random.default_rng(seed)
    halt_final = len(program) + 8
    labels = (halt_final, 3, 0x4000, 0x5000)
    log = []
    sponge_ops = list(sponge_ops or [])

    def keccak(addr_word, length, clock):
        # KECCAK_GENERAL (cpu_stark.rs:33-56): the Cpu row sends (context, segment, virt, len, (clock - 1) * NUM_CHANNELS + 1, digest)
        virt, seg, ctx = [(addr_word >> (32 * i)) & 0xFFFFFFFF for i in range(3)]
        data = keccak_inputs[(ctx, seg, virt)]
        assert len(data) == length
        op = (ctx, seg, virt, (clock - 1) * num_channels + 1, data)
        sponge_ops.append(op)
        return int.from_bytes(keccak_sponge_trace(8, [op])[1][0], "big")
    packing_ops = list(packing_ops or [])
    code_bytes = {}                                 # contents of the (pre-initialised) code segments the 32-byte loads read

    def load32(addr_word, length, clock):
        # MLOAD_32BYTES (cpu_stark.rs:150-172 ctl_data_byte_packing): the Cpu row sends (1, context, segment, virt, len, (clock - 1) * NUM_CHANNELS + 1,
        # packed value); byte k of the sequence lives at virt + k, the BytePacking row holds them least significant first
        virt, seg, ctx = [(addr_word >> (32 * i)) & 0xFFFFFFFF for i in range(3)]
        assert seg == 0, "reads of unwritten memory are only free in a pre-initialised segment (Segment::Code)"
        data = bytes(code_bytes.setdefault((ctx, seg, virt + k), int(rng.integers(0, 256))) for k in range(length))
        packing_ops.append((1, ctx, seg, virt, (clock - 1) * num_channels + 1, data[::-1]))
        return data

    def store32(addr_word, value, ln, clock):
        # MSTORE_32BYTES_n (cpu_stark.rs:174-223 ctl_data_byte_unpacking): (0, context, segment, virt, len = new offset - virt, timestamp, value)
        virt, seg, ctx = [(addr_word >> (32 * i)) & 0xFFFFFFFF for i in range(3)]
        packing_ops.append((0, ctx, seg, virt, (clock - 1) * num_channels + 1, value.to_bytes(ln, "little")))
    def push32(code_ctx, virt, imm, clock):
        # a user-mode PUSH (cpu_stark.rs:264-304 ctl_data_byte_packing_push): (1, code context, Segment::Code, pc + 1, len, timestamp, pushed word)
        packing_ops.append((1, code_ctx, 0, virt, (clock - 1) * num_channels + 1, imm[::-1]))
    stale = []                                      # contexts pruned by SET_CONTEXT rows: the Memory table lists them (context pruning lookup)

    def jumptable(virt, handler, clock):
        # syscall / exception rows (cpu_stark.rs:225-262 ctl_data_jumptable_read): (1, 0, Segment::Code, virt, 3, timestamp, handler address);
        # the three bytes are the kernel's jump-table entry, big-endian (a pre-initialised segment: any content is admissible)
        packing_ops.append((1, 0, 0, virt, (clock - 1) * num_channels + 1, handler.to_bytes(3, "little")))
    cpu = cpu_program_trace(log_cpu, program, halt_final=halt_final, log=log, inputs=inputs, keccak=keccak if keccak_inputs is not None else None,
                            load32=load32, store32=store32, syscall_jumptable=labels[2], exception_jumptable=labels[3], jumptable=jumptable,
                            stale=stale, push32=push32)
    limbs = lambda x: [(x >> (32 * i)) & 0xFFFFFFFF for i in range(8)]
    NUM_CHANNELS = num_channels                     # 5 in the reference; the parameter exists for the negative test
    ops = []                                        # (ctx, seg, virt, timestamp, is_read, filter, value limbs)
    for r in range(cpu.shape[1]):
        row = [int(v) for v in cpu[:, r]]
        if sum(row[6:24]) == 0:
            continue                                # halt rows send nothing
        ts = lambda channel: row[40] * NUM_CHANNELS + channel - NUM_CHANNELS + 1
        opcode = sum(row[24 + b] << b for b in range(8))
        ops.append((row[1], 0, row[2], ts(0), 1, 1, [opcode] + [0] * 7))                    # code read: code_context, Segment::Code, pc
        for c in range(3):
            o = 41 + 13 * c
            if row[o]:
                ops.append((row[o + 2], row[o + 3], row[o + 4], ts(1 + c), row[o + 1], 1, row[o + 5:o + 13]))
        if row[80]:
            ops.append((row[82], row[83], row[84], ts(4), row[81], 1, row[46:54]))          # partial channel: the value of mem_channels[0]
        if row[17] and row[24]:
            # SET_CONTEXT (cpu_stark.rs:391-431): the old stack pointer (stack_len - 1) is written to the old context's
            # ContextMetadata::StackSize cell at the time of channel 2, the new one (next row's stack_len) read from the new context's at channel 3
            ops.append((row[0], 6, 11, ts(2), 0, 1, [row[3] - 1] + [0] * 7))
            ops.append((row[48], 6, 11, ts(3), 1, 1, [int(cpu[3, r + 1])] + [0] * 7))
    logic_ops = []
    sponge = keccak = None
    if sponge_ops:
        # KeccakSponge operations that no Cpu row asked for (lookup 2, Cpu -> KeccakSponge, stays unbalanced): the sponge rows send
        # their permutations to the Keccak table (lookups 3, 4), their xors to the Logic table (5), their input bytes to Memory (6)
        sponge, _ = keccak_sponge_trace(8, sponge_ops)
        nrows = sum(len(o[4]) // 136 + 1 for o in sponge_ops)
        lanes, stamps = [], []
        for i in range(nrows):
            row = [int(v) for v in sponge[:, i]]
            st32 = row[328:362] + row[176:192]                                              # xored rate ++ original capacity
            lanes.append([st32[2 * w] | (st32[2 * w + 1] << 32) for w in range(25)])
            stamps.append(row[4])
            block32 = [int.from_bytes(bytes(row[192 + 4 * w:196 + 4 * w]), "little") for w in range(34)]
            for c in range(5):                                                              # num_logic_ctls (keccak_sponge_stark.rs:137-186)
                in0 = (row[142:176] + [0] * 6)[8 * c:8 * c + 8]
                in1 = (block32 + [0] * 6)[8 * c:8 * c + 8]
                pack = lambda w: [w[2 * l] | (w[2 * l + 1] << 32) for l in range(4)]
                logic_ops.append([2] + pack(in0) + pack(in1))
            nread = 136 if row[0] else row[6:142].index(1)                                  # full block, or the input bytes of the final block
            for b in range(nread):
                ops.append((row[1], row[2], row[3] + row[5] + b, row[4], 1, 1, [row[192 + b]] + [0] * 7))   # ctl_looking_memory(b), :106-134
        keccak, _ = keccak_trace(max(5, (24 * nrows - 1).bit_length()), np.array(lanes, dtype=np.uint64), np.array(stamps, dtype=np.uint64))
    packing = None
    if packing_ops:
        # BytePacking operations: the ones the Cpu's MLOAD_32BYTES / MSTORE_32BYTES rows ask for (load32 / store32 above: lookup 1, Cpu ->
        # BytePacking, balances for these) and hand-placed ones nobody asked for (`packing_ops` argument: lookup 1 stays open): byte i of a
        # sequence of length L lives at virt + L - 1 - i (byte_packing_stark.rs:105-148)
        packing = np.zeros((71, 256), dtype=np.uint64)
        for j, (is_read, ctx, seg, virt, t_, data) in enumerate(packing_ops):
            L = len(data)
            assert 1 <= L <= 32
            packing[0, j], packing[L, j] = is_read, 1                                        # is_read, index_len[L - 1]
            packing[33:37, j] = [ctx, seg, virt, t_]
            packing[37:37 + L, j] = list(data)
            for i in range(L):
                ops.append((ctx, seg, virt + L - 1 - i, t_, is_read, 1, [data[i]] + [0] * 7))
        packing[69] = np.minimum(np.arange(256), 255).astype(np.uint64)
        packing[70, :256] = np.bincount(packing[37:69].astype(np.int64).ravel(), minlength=256).astype(np.uint64)
    for row in extra_memory_rows:                                                           # (is_read, context, segment, virt, value limbs, timestamp)
        ops.append((row[1], row[2], row[3], row[12], row[0], 1, [int(x) for x in row[4:12]]))
    before_addrs = [(0, 5, 100 + i) for i in range(k_before)]
    before_vals = rng.integers(1, 1 << 32, size=(k_before, 8), dtype=np.uint64)
    for a, v in zip(before_addrs, before_vals):
        ops.append((a[0], a[1], a[2], 0, 0, 1, [int(x) for x in v]))                        # memory_stark.rs:408-423
    ops.sort(key=lambda o: o[:4])
    n = 1 << log_mem
    if ops[0][:3] != (0, 0, 0):                                                             # fill_gaps: dummy read at (0, 0, 0)
        ops.insert(0, (0, 0, 0, 1, 1, 0, [0] * 8))
    last = ops[-1]
    assert len(ops) < n
    while len(ops) < n:                                                                     # pad_memory_ops
        ops.append((last[0], last[1], last[2] + 1, last[3] + 1, 1, 0, [0] * 8))
    m = np.zeros((14, n), dtype=np.uint64)
    for i, (ctx, seg, virt, t_, is_read, filt, val) in enumerate(ops):
        m[:6, i] = [filt, t_, is_read, ctx, seg, virt]
        m[6:, i] = val
    memory = memory_finish_reference(m, stale=stale)
    after = [(int(memory[4, i]), int(memory[5, i]), int(memory[6, i])) for i in range(n) if memory[26, i]]
    after_vals = [[int(memory[7 + l, i]) for l in range(8)] for i in range(n) if memory[26, i]]
    # Arithmetic: the rows of the executed operations, then padding and the range-check columns
    arith = np.zeros((116, 1 << 16), dtype=np.uint64)
    r = 0
    names = {"D": "div", "O": "mod", "a": "addmod", "m": "mulmod"}
    for e in log:
        if e[0] == "M":
            out, lo, hi = arithmetic_mul_rows(e[1], e[2])
            arith[1, r] = 1
            arith[18:34, r], arith[34:50, r] = _limbs(e[1]), _limbs(e[2])
            arith[66:82, r], arith[82:98, r], arith[98:114, r] = out, lo, hi
            r += 1
        elif e[0] == "I":                              # PROVER_INPUT is range-checked by the Arithmetic table (arithmetic/mod.rs:343-359)
            arith[16, r], arith[17, r] = 1, 0xee       # IS_RANGE_CHECK, OPCODE_COL
            arith[18:34, r], arith[66:82, r] = _limbs(e[1]), _limbs(e[2])
            r += 1
        elif e[0] in "Yx":                             # syscall / exception rows are range-checked too (operation.rs:777-790): opcode, stack top, handler, 0 -> kexit_info
            arith[16, r], arith[17, r] = 1, e[1]
            arith[18:34, r], arith[34:50, r], arith[66:82, r] = _limbs(e[2]), _limbs(e[3]), _limbs(e[4])
            r += 1
        elif e[0] in "ASLG":                           # addcy.rs: ADD in0 + in1 = out + cy 2^256; SUB / LT / GT by rearranging it
            a, b, M = e[1], e[2], 1 << 256
            if e[0] == "A":
                out, aux = (a + b) % M, (a + b) >> 256
            elif e[0] == "S":
                out, aux = (a - b) % M, int(a < b)
            elif e[0] == "L":
                out, aux = int(a < b), (a - b) % M
            else:
                out, aux = int(a > b), (b - a) % M
            assert out == e[-1]
            arith[{"A": 0, "S": 2, "L": 11, "G": 12}[e[0]], r] = 1
            arith[18:34, r], arith[34:50, r], arith[66:82, r], arith[82:98, r] = _limbs(a), _limbs(b), _limbs(out), _limbs(aux)
            r += 1
        elif e[0] in names:
            row1, row2, res = arithmetic_modular_rows(names[e[0]], e[1], e[2], e[3] if e[0] in "am" else 0)
            assert res == e[-1]
            arith[:, r], arith[:, r + 1] = row1, row2
            r += 2
        elif e[0] in "&|^":
            logic_ops.append(["&|^".index(e[0])] + [(e[1] >> (64 * l)) & 0xFFFFFFFFFFFFFFFF for l in range(4)]
                             + [(e[2] >> (64 * l)) & 0xFFFFFFFFFFFFFFFF for l in range(4)])
        else:
            raise ValueError("instruction %r needs a table this segment does not build" % e[0])
    arith[114] = np.minimum(np.arange(1 << 16), 65535).astype(np.uint64)
    arith[115, :65536] = np.bincount(arith[18:114].astype(np.int64).ravel(), minlength=65536).astype(np.uint64)
    tr = [None] * 9
    tr[T_ARITHMETIC], tr[T_CPU], tr[T_MEMORY] = arith, cpu, memory
    if logic_ops:
        tr[T_LOGIC] = logic_trace_from_ops(max(log_logic, (len(logic_ops) - 1).bit_length()), np.array(logic_ops, dtype=np.uint64))
    tr[T_KECCAK_SPONGE], tr[T_KECCAK], tr[T_BYTE_PACKING] = sponge, keccak, packing
    tr[T_MEM_BEFORE] = memcont_trace_from(log_memcont, before_addrs, before_vals)
    tr[T_MEM_AFTER] = memcont_trace_from(max(log_memcont, (len(after) - 1).bit_length()), after, after_vals)
    return tr, labels


CPU_SEGMENT_PROGRAM = ("0PPPSuAuAPAiNJ" "X" "PPvwstuE" "PP&PP|^" "PPMPPDPPOPPPaPPPm" "PPLPPGNZC" "XXXXXXXXXXXXJ")     # 69 instructions, 68 executed


def random_segment(log_ns, seed):
    """uniformly random traces of the given heights (log_ns[table] or None) — throughput / parity inputs, not valid witnesses"""
    return [None if lg is None else random_trace(t, lg, seed * 16 + t) for t, lg in enumerate(log_ns)]
