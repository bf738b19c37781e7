// CpuStark: constraints (in the emission order of cpu_stark.rs:594-626) and cross-table-lookup columns.
// Source: /root/reference/evm_arithmetization/src/cpu/{columns/mod.rs:56-97, columns/ops.rs:6-47, columns/general.rs:11-246,
// cpu_stark.rs:33-470 (CTL), byte_unpacking.rs:11, clock.rs:15, contextops.rs:40-330, control_flow.rs:50-104, decode.rs:86-220,
// dup_swap.rs, gas.rs, halt.rs, jumps.rs, membus.rs, memio.rs, modfp254.rs, pc.rs, push0.rs, shift.rs, simple_logic/*, stack.rs,
// syscalls_exceptions.rs}.  Default feature set (eth_mainnet): 85 columns.
#pragma once
#include "hd.h"
#include "lookup.h"
#include "consumer.h"

namespace zkstark { namespace cpu {

enum : uint32_t {
    CONTEXT = 0, CODE_CONTEXT = 1, PROGRAM_COUNTER = 2, STACK_LEN = 3, IS_KERNEL_MODE = 4, GAS = 5,
    // OpsColumnsView (columns/ops.rs), iteration order of `lv.op` / COL_MAP.op
    OP_BEGIN = 6,
    OP_BINARY_OP = 6, OP_TERNARY_OP = 7, OP_FP254_OP = 8, OP_EQ_ISZERO = 9, OP_LOGIC_OP = 10, OP_NOT_POP = 11, OP_SHIFT = 12,
    OP_JUMPDEST_KECCAK_GENERAL = 13, OP_JUMPS = 14, OP_PUSH_PROVER_INPUT = 15, OP_DUP_SWAP = 16, OP_CONTEXT_OP = 17,
    OP_M_OP_32BYTES = 18, OP_EXIT_KERNEL = 19, OP_M_OP_GENERAL = 20, OP_PC_PUSH0 = 21, OP_SYSCALL = 22, OP_EXCEPTION = 23,
    OP_END = 24, NUM_OPS = 18,
    OPCODE_BITS = 24,    // 8, little endian
    GENERAL = 32,        // 8 shared columns (union)
    CLOCK = 40,
    MEM_CHANNELS = 41,   // 3 channels x 13
    PARTIAL_CHANNEL = 80,
    NUM_COLUMNS = 85
};
// NUM_CHANNELS = channel_indices::GP.end + 1 (membus.rs:11-39): the code channel, the three general-purpose channels and the partial channel
static const uint32_t NUM_GP_CHANNELS = 3, NUM_CHANNELS = 5, VALUE_LIMBS = 8, CHANNEL_WIDTH = 13;
// general-column views (columns/general.rs)
enum : uint32_t {
    G_EXC_CODE_BITS = GENERAL,          // exception: 3
    G_LOGIC_DIFF_PINV = GENERAL,        // logic: 8
    G_JUMPS_SHOULD_JUMP = GENERAL, G_JUMPS_COND_SUM_PINV = GENERAL + 1,
    G_SHIFT_HIGH_LIMB_SUM_INV = GENERAL,
    G_STACK_INV = GENERAL + 4, G_STACK_INV_AUX = GENERAL + 5, G_STACK_INV_AUX_2 = GENERAL + 6, G_STACK_LEN_BOUNDS_AUX = GENERAL + 7,
    G_PUSH_IS_NOT_KERNEL = GENERAL,
    G_PRUNING_FLAG = GENERAL
};
ZKS_HD uint32_t ch_used(uint32_t c) { return MEM_CHANNELS + CHANNEL_WIDTH * c; }
ZKS_HD uint32_t ch_is_read(uint32_t c) { return MEM_CHANNELS + CHANNEL_WIDTH * c + 1; }
ZKS_HD uint32_t ch_addr_context(uint32_t c) { return MEM_CHANNELS + CHANNEL_WIDTH * c + 2; }
ZKS_HD uint32_t ch_addr_segment(uint32_t c) { return MEM_CHANNELS + CHANNEL_WIDTH * c + 3; }
ZKS_HD uint32_t ch_addr_virtual(uint32_t c) { return MEM_CHANNELS + CHANNEL_WIDTH * c + 4; }
ZKS_HD uint32_t ch_value(uint32_t c, uint32_t limb) { return MEM_CHANNELS + CHANNEL_WIDTH * c + 5 + limb; }
enum : uint32_t { PC_USED = 80, PC_IS_READ = 81, PC_ADDR_CONTEXT = 82, PC_ADDR_SEGMENT = 83, PC_ADDR_VIRTUAL = 84 };

// memory/segments.rs:14-88 (unscaled segment numbers) and kernel/constants/context_metadata.rs
enum : uint64_t { SEG_CODE = 0, SEG_STACK = 1, SEG_MAIN_MEMORY = 2, SEG_CALLDATA = 3, SEG_RETURNDATA = 4, SEG_GLOBAL_METADATA = 5,
                  SEG_CONTEXT_METADATA = 6, SEG_KERNEL_GENERAL = 7, SEG_JUMPDEST_BITS = 14, SEG_REGISTERS_STATES = 33 };
static const unsigned SEGMENT_SCALING_FACTOR = 32;

// ---- byte_unpacking.rs:11-44 -------------------------------------------------------------------------------------
template <class P, class V, class CC>
ZKS_HD void eval_byte_unpacking(const V& lv, const V& nv, CC& yc) {
    // MSTORE_32BYTES is differentiated from MLOAD_32BYTES by the 5th bit set to 0
    P filter = lv[OP_M_OP_32BYTES] * (lv[OPCODE_BITS + 5] - P::one());
    P len_bits = P::zero();
    for (uint32_t i = 0; i < 5; i++) len_bits = len_bits + lv[OPCODE_BITS + i] * P::from_u64(1ULL << i);
    P len = len_bits + P::one();
    yc.constraint(filter * (nv[ch_value(0, 0)] - lv[ch_value(0, 0)] - len));
    yc.constraint(filter * (nv[ch_value(0, 1)] - lv[ch_value(0, 1)]));
    yc.constraint(filter * (nv[ch_value(0, 2)] - lv[ch_value(0, 2)]));
    for (uint32_t i = 3; i < VALUE_LIMBS; i++) yc.constraint(filter * nv[ch_value(0, i)]);
}

// ---- clock.rs:15-25 -------------------------------------------------------------------------------------------------
template <class P, class V, class CC>
ZKS_HD void eval_clock(const V& lv, const V& nv, CC& yc) {
    yc.constraint_first_row(lv[CLOCK] - P::one());
    yc.constraint_transition(nv[CLOCK] - lv[CLOCK] - P::one());
}

// ---- contextops.rs ---------------------------------------------------------------------------------------------------
template <class P, class V, class CC>
ZKS_HD void eval_contextops(const V& lv, const V& nv, CC& yc) {
    const P one = P::one();
    // eval_packed_keep: every op but context_op keeps the context (KEEPS_CONTEXT, contextops.rs:17-38)
    for (uint32_t op = OP_BEGIN; op < OP_END; op++)
        if (op != OP_CONTEXT_OP) yc.constraint_transition(lv[op] * (nv[CONTEXT] - lv[CONTEXT]));
    {
        P is_get_context = lv[OP_CONTEXT_OP] * (lv[OPCODE_BITS] - one);
        yc.constraint_transition(is_get_context * (nv[CONTEXT] - lv[CONTEXT]));
    }
    // eval_packed_get
    {
        P filter = lv[OP_CONTEXT_OP] * (one - lv[OPCODE_BITS]);
        yc.constraint(filter * (nv[ch_value(0, 2)] - lv[CONTEXT]));
        for (uint32_t i = 0; i < VALUE_LIMBS; i++) if (i != 2) yc.constraint(filter * nv[ch_value(0, i)]);
        yc.constraint(filter * lv[G_PRUNING_FLAG]);
        yc.constraint(filter * (nv[STACK_LEN] - (lv[STACK_LEN] + one)));
        yc.constraint(filter * lv[ch_used(1)]);          // disable_unused_channels(lv, filter, vec![1])
        yc.constraint(filter * nv[ch_used(0)]);
    }
    // eval_packed_set
    {
        P filter = lv[OP_CONTEXT_OP] * lv[OPCODE_BITS];
        yc.constraint(filter * (lv[ch_value(0, 2)] - nv[CONTEXT]));
        // stack_top[1..] except (relative) index 1, i.e. limbs 1, 3, 4, 5, 6, 7
        for (uint32_t i = 1; i < VALUE_LIMBS; i++) if (i != 2) yc.constraint(filter * lv[ch_value(0, i)]);
        yc.constraint(lv[OP_CONTEXT_OP] * lv[G_PRUNING_FLAG] * (lv[G_PRUNING_FLAG] - one));
        yc.constraint(filter * (lv[G_PRUNING_FLAG] - lv[ch_value(0, 0)]));
        yc.constraint(lv[OP_CONTEXT_OP] * (lv[G_STACK_INV_AUX] * lv[OPCODE_BITS] - lv[G_STACK_INV_AUX_2]));
        for (uint32_t i = 0; i < VALUE_LIMBS; i++)
            yc.constraint(lv[OP_CONTEXT_OP] * lv[G_STACK_INV_AUX_2] * (nv[ch_value(0, i)] - lv[ch_value(2, i)]));
        yc.constraint(filter * lv[ch_used(1)]);          // disable_unused_channels(lv, filter, vec![1])
        yc.constraint(filter * nv[ch_used(0)]);
    }
    // stack constraints shared by GET_CONTEXT and SET_CONTEXT (memory channel 2)
    {
        P filter = lv[OP_CONTEXT_OP];
        P stack_len = nv[STACK_LEN] - (one - lv[OPCODE_BITS]);
        yc.constraint(filter * (stack_len * lv[G_STACK_INV] - lv[G_STACK_INV_AUX]));
        yc.constraint(filter * (lv[G_STACK_INV_AUX] - lv[ch_used(2)]));
        P new_filter = filter * lv[G_STACK_INV_AUX];
        yc.constraint(new_filter * (lv[ch_is_read(2)] - lv[OPCODE_BITS]));
        yc.constraint(new_filter * (lv[ch_addr_context(2)] - nv[CONTEXT]));
        yc.constraint(new_filter * (lv[ch_addr_segment(2)] - P::from_u64(SEG_STACK)));
        P addr_virtual = stack_len - one;
        yc.constraint(new_filter * (lv[ch_addr_virtual(2)] - addr_virtual));
    }
}

// ---- control_flow.rs:50-104 -------------------------------------------------------------------------------------------
template <class P, class V, class CC>
ZKS_HD void eval_control_flow(const V& lv, const V& nv, CC& yc, const TableParams& prm) {
    const P one = P::one();
    P is_cpu_cycle = P::zero(), is_cpu_cycle_next = P::zero();
    for (uint32_t op = OP_BEGIN; op < OP_END; op++) { is_cpu_cycle = is_cpu_cycle + lv[op]; is_cpu_cycle_next = is_cpu_cycle_next + nv[op]; }
    P next_halt_state = one - is_cpu_cycle_next;
    yc.constraint_transition(is_cpu_cycle * (is_cpu_cycle_next + next_halt_state - one));
    // NATIVE_INSTRUCTIONS (control_flow.rs:13-34)
    P is_native = lv[OP_BINARY_OP] + lv[OP_TERNARY_OP] + lv[OP_FP254_OP] + lv[OP_EQ_ISZERO] + lv[OP_LOGIC_OP] + lv[OP_NOT_POP] +
                  lv[OP_SHIFT] + lv[OP_JUMPDEST_KECCAK_GENERAL] + lv[OP_PC_PUSH0] + lv[OP_DUP_SWAP] + lv[OP_CONTEXT_OP] +
                  lv[OP_M_OP_GENERAL];
    yc.constraint_transition(is_native * (lv[PROGRAM_COUNTER] - nv[PROGRAM_COUNTER] + one));
    yc.constraint_transition(is_native * (lv[IS_KERNEL_MODE] - nv[IS_KERNEL_MODE]));
    P is_prover_input = lv[OP_PUSH_PROVER_INPUT] * lv[OPCODE_BITS + 7];
    yc.constraint_transition(is_prover_input * (lv[PROGRAM_COUNTER] - nv[PROGRAM_COUNTER] + one));
    yc.constraint_transition(is_prover_input * (lv[IS_KERNEL_MODE] - nv[IS_KERNEL_MODE]));
    yc.constraint(lv[OP_PUSH_PROVER_INPUT] * ((lv[IS_KERNEL_MODE] + lv[G_PUSH_IS_NOT_KERNEL]) - one));
    P is_last_noncpu_cycle = (is_cpu_cycle - one) * is_cpu_cycle_next;
    P pc_diff = nv[PROGRAM_COUNTER] - P::from_u64(prm.init);
    yc.constraint_transition(is_last_noncpu_cycle * pc_diff);
    yc.constraint_transition(is_last_noncpu_cycle * (nv[IS_KERNEL_MODE] - one));
    yc.constraint_transition(is_last_noncpu_cycle * nv[STACK_LEN]);
}

// ---- decode.rs:86-220 ---------------------------------------------------------------------------------------------------
template <class P, class V, class CC>
ZKS_HD void eval_decode(const V& lv, CC& yc) {
    const P one = P::one();
    // OPCODES: (first opcode, block length log2, kernel only, flag column)   decode.rs:34-52
    const uint32_t OPC_OC[5] = {0x14, 0x56, 0x80, 0xf6, 0xf9};
    const uint32_t OPC_BL[5] = {1, 1, 5, 1, 0};
    const bool OPC_KO[5] = {false, false, false, true, true};
    const uint32_t OPC_COL[5] = {OP_EQ_ISZERO, OP_JUMPS, OP_DUP_SWAP, OP_CONTEXT_OP, OP_EXIT_KERNEL};
    const uint32_t COMBINED[11] = {OP_LOGIC_OP, OP_FP254_OP, OP_BINARY_OP, OP_TERNARY_OP, OP_SHIFT, OP_M_OP_GENERAL,
                                   OP_JUMPDEST_KECCAK_GENERAL, OP_NOT_POP, OP_PC_PUSH0, OP_M_OP_32BYTES, OP_PUSH_PROVER_INPUT};
    P kernel_mode = lv[IS_KERNEL_MODE];
    yc.constraint(kernel_mode * (kernel_mode - one));
    for (uint32_t i = 0; i < 8; i++) { P bit = lv[OPCODE_BITS + i]; yc.constraint(bit * (bit - one)); }
    for (uint32_t k = 0; k < 5; k++) { P flag = lv[OPC_COL[k]]; yc.constraint(flag * (flag - one)); }
    for (uint32_t k = 0; k < 11; k++) { P flag = lv[COMBINED[k]]; yc.constraint(flag * (flag - one)); }
    P flag_sum = P::zero();
    for (uint32_t k = 0; k < 5; k++) flag_sum = flag_sum + lv[OPC_COL[k]];
    for (uint32_t k = 0; k < 11; k++) flag_sum = flag_sum + lv[COMBINED[k]];
    yc.constraint(flag_sum * (flag_sum - one));
    for (uint32_t k = 0; k < 5; k++) {
        P unavailable = OPC_KO[k] ? one - kernel_mode : P::zero();
        P mismatch = P::zero();
        // the top (8 - block_length) bits, from bit 7 downwards
        for (uint32_t t = 0; t < 8 - OPC_BL[k]; t++) {
            uint32_t i = 7 - t;
            P row_bit = lv[OPCODE_BITS + i];
            mismatch = mismatch + (((OPC_OC[k] >> i) & 1) ? one - row_bit : row_bit);
        }
        yc.constraint(lv[OPC_COL[k]] * (unavailable + mismatch));
    }
    P opcode = P::zero(), opcode_high_three = P::zero();
    for (uint32_t i = 0; i < 8; i++) {
        P term = lv[OPCODE_BITS + i] * P::from_u64(1ULL << i);
        opcode = opcode + term;
        if (i >= 5) opcode_high_three = opcode_high_three + term;
    }
    yc.constraint((kernel_mode - one) * lv[OP_FP254_OP]);
    yc.constraint(lv[OP_TERNARY_OP] * lv[OPCODE_BITS + 1] * (kernel_mode - one));
    yc.constraint((kernel_mode - one) * lv[OP_M_OP_GENERAL]);
    yc.constraint((opcode - P::from_u64(0xfb)) * (opcode - P::from_u64(0xfc)) * lv[OP_M_OP_GENERAL]);
    yc.constraint((kernel_mode - one) * lv[OP_JUMPDEST_KECCAK_GENERAL] * (one - lv[OPCODE_BITS + 1]));
    yc.constraint((opcode - P::from_u64(0x21)) * (opcode - P::from_u64(0x5b)) * lv[OP_JUMPDEST_KECCAK_GENERAL]);
    yc.constraint((opcode - P::from_u64(0x58)) * (opcode - P::from_u64(0x5f)) * lv[OP_PC_PUSH0]);
    yc.constraint((opcode - P::from_u64(0x19)) * (opcode - P::from_u64(0x50)) * lv[OP_NOT_POP]);
    yc.constraint((kernel_mode - one) * lv[OP_M_OP_32BYTES]);
    yc.constraint((opcode_high_three - P::from_u64(0xc0)) * (opcode - P::from_u64(0xf8)) * lv[OP_M_OP_32BYTES]);
    yc.constraint((opcode - P::from_u64(0xee)) * (opcode_high_three - P::from_u64(0x60)) * lv[OP_PUSH_PROVER_INPUT]);
    yc.constraint(lv[OP_PUSH_PROVER_INPUT] * lv[OPCODE_BITS + 7] * (kernel_mode - one));
}

// ---- stack.rs:193-316 eval_packed_one ------------------------------------------------------------------------------
struct StackBehavior { uint32_t num_pops; bool pushes; bool disable_other_channels; };

template <class P, class V, class CC>
ZKS_HD void eval_stack_one(const V& lv, const V& nv, P filter, StackBehavior sb, CC& yc) {
    const P one = P::one();
    const P seg_stack = P::from_u64(SEG_STACK);
    if (sb.num_pops > 0) {
        for (uint32_t i = 1; i < sb.num_pops; i++) {
            yc.constraint(filter * (lv[ch_used(i)] - one));
            yc.constraint(filter * (lv[ch_is_read(i)] - one));
            yc.constraint(filter * (lv[ch_addr_context(i)] - lv[CONTEXT]));
            yc.constraint(filter * (lv[ch_addr_segment(i)] - seg_stack));
            P addr_virtual = lv[STACK_LEN] - P::from_u64(i + 1);
            yc.constraint(filter * (lv[ch_addr_virtual(i)] - addr_virtual));
        }
        yc.constraint(filter * lv[PC_USED]);
        if (!sb.pushes) {
            P len_diff = lv[STACK_LEN] - P::from_u64(sb.num_pops);
            P new_filter = len_diff * filter;
            yc.constraint_transition(new_filter * (nv[ch_used(0)] - one));
            yc.constraint_transition(new_filter * (nv[ch_is_read(0)] - one));
            yc.constraint_transition(new_filter * (nv[ch_addr_context(0)] - nv[CONTEXT]));
            yc.constraint_transition(new_filter * (nv[ch_addr_segment(0)] - seg_stack));
            P addr_virtual = nv[STACK_LEN] - one;
            yc.constraint_transition(new_filter * (nv[ch_addr_virtual(0)] - addr_virtual));
            yc.constraint(filter * (len_diff * lv[G_STACK_INV] - lv[G_STACK_INV_AUX]));
            P empty_stack_filter = filter * (lv[G_STACK_INV_AUX] - one);
            yc.constraint_transition(empty_stack_filter * nv[ch_used(0)]);
        }
    } else if (sb.pushes) {
        P new_filter = lv[STACK_LEN] * filter;
        yc.constraint(new_filter * (lv[PC_USED] - one));
        yc.constraint(new_filter * lv[PC_IS_READ]);
        yc.constraint(new_filter * (lv[PC_ADDR_CONTEXT] - lv[CONTEXT]));
        yc.constraint(new_filter * (lv[PC_ADDR_SEGMENT] - seg_stack));
        P addr_virtual = lv[STACK_LEN] - one;
        yc.constraint(new_filter * (lv[PC_ADDR_VIRTUAL] - addr_virtual));
        yc.constraint(filter * (lv[STACK_LEN] * lv[G_STACK_INV] - lv[G_STACK_INV_AUX]));
        P empty_stack_filter = filter * (lv[G_STACK_INV_AUX] - one);
        yc.constraint(empty_stack_filter * lv[PC_USED]);
    } else {
        yc.constraint(filter * nv[ch_used(0)]);
        for (uint32_t i = 0; i < VALUE_LIMBS; i++) yc.constraint(filter * (lv[ch_value(0, i)] - nv[ch_value(0, i)]));
        yc.constraint(filter * lv[PC_USED]);
    }
    if (sb.disable_other_channels) {
        uint32_t lo = sb.num_pops > 1 ? sb.num_pops : 1, hi = NUM_GP_CHANNELS - (sb.pushes ? 1 : 0);
        for (uint32_t i = lo; i < hi; i++) yc.constraint(filter * lv[ch_used(i)]);
    }
    P num_pops = P::from_u64(sb.num_pops), push = P::from_u64(sb.pushes ? 1 : 0);
    yc.constraint_transition(filter * (nv[STACK_LEN] - (lv[STACK_LEN] - num_pops + push)));
}

// ---- dup_swap.rs ---------------------------------------------------------------------------------------------------
// channels_equal_packed: `a` row channel ca == `b` row channel cb
template <class P, class V, class CC>
ZKS_HD void channels_equal(P filter, const V& a, uint32_t ca, const V& b, uint32_t cb, CC& yc) {
    for (uint32_t i = 0; i < VALUE_LIMBS; i++) yc.constraint(filter * (a[ch_value(ca, i)] - b[ch_value(cb, i)]));
}
template <class P, class V, class CC>
ZKS_HD void constrain_channel(bool is_read, P filter, P offset, uint32_t ch, const V& lv, CC& yc) {
    const P one = P::one();
    yc.constraint(filter * (lv[ch_used(ch)] - one));
    yc.constraint(filter * (lv[ch_is_read(ch)] - P::from_u64(is_read ? 1 : 0)));
    yc.constraint(filter * (lv[ch_addr_context(ch)] - lv[CONTEXT]));
    yc.constraint(filter * (lv[ch_addr_segment(ch)] - P::from_u64(SEG_STACK)));
    P addr_virtual = lv[STACK_LEN] - one - offset;
    yc.constraint(filter * (lv[ch_addr_virtual(ch)] - addr_virtual));
}
template <class P, class V, class CC>
ZKS_HD void eval_dup_swap(const V& lv, const V& nv, CC& yc) {
    const P one = P::one();
    P n = lv[OPCODE_BITS] + lv[OPCODE_BITS + 1] * P::from_u64(2) + lv[OPCODE_BITS + 2] * P::from_u64(4) + lv[OPCODE_BITS + 3] * P::from_u64(8);
    {   // DUP: write channel 1, read channel 2
        P filter = lv[OP_DUP_SWAP] * (one - lv[OPCODE_BITS + 4]);
        channels_equal<P>(filter, lv, 1, lv, 0, yc);
        constrain_channel<P>(false, filter, P::zero(), 1, lv, yc);
        channels_equal<P>(filter, lv, 2, nv, 0, yc);
        constrain_channel<P>(true, filter, n, 2, lv, yc);
        yc.constraint_transition(filter * (nv[STACK_LEN] - lv[STACK_LEN] - one));
        yc.constraint(filter * nv[ch_used(0)]);
    }
    {   // SWAP: in1 = channel 0, in2 = channel 1, out = channel 2
        P n_plus_one = n + one;
        P filter = lv[OP_DUP_SWAP] * lv[OPCODE_BITS + 4];
        channels_equal<P>(filter, lv, 0, lv, 2, yc);
        constrain_channel<P>(false, filter, n_plus_one, 2, lv, yc);
        channels_equal<P>(filter, lv, 1, nv, 0, yc);
        constrain_channel<P>(true, filter, n_plus_one, 1, lv, yc);
        yc.constraint(filter * (nv[STACK_LEN] - lv[STACK_LEN]));
        yc.constraint(filter * nv[ch_used(0)]);
    }
    yc.constraint(lv[OP_DUP_SWAP] * lv[PC_USED]);
}

// ---- gas.rs ----------------------------------------------------------------------------------------------------------
template <class P, class V, class CC>
ZKS_HD void eval_gas(const V& lv, const V& nv, CC& yc) {
    const P one = P::one();
    // SIMPLE_OPCODES in OpsColumnsView field order (gas.rs:21-42); -1 = None
    const int COST[NUM_OPS] = {-1, -1, 0, 3, 3, -1, 3, -1, -1, -1, 3, 0, 0, -1, 0, 2, -1, -1};
    const uint64_t G_JUMPDEST = 1, G_BASE = 2, G_VERYLOW = 3, G_LOW = 5, G_MID = 8, G_HIGH = 10, KERNEL_ONLY = 0;
    P filter = P::zero(), gas_used = P::zero();
    for (uint32_t i = 0; i < NUM_OPS; i++)
        if (COST[i] >= 0) { filter = filter + lv[OP_BEGIN + i]; gas_used = gas_used + P::from_u64((uint64_t)COST[i]) * lv[OP_BEGIN + i]; }
    yc.constraint_transition(filter * (nv[GAS] - (lv[GAS] + gas_used)));
    P gas_diff = nv[GAS] - lv[GAS];
    for (uint32_t i = 0; i < NUM_OPS; i++)
        if (COST[i] >= 0) yc.constraint_transition(lv[OP_BEGIN + i] * (gas_diff - P::from_u64((uint64_t)COST[i])));
    P jump_gas_cost = P::from_u64(G_MID) + lv[OPCODE_BITS] * P::from_u64(G_HIGH - G_MID);
    yc.constraint_transition(lv[OP_JUMPS] * (gas_diff - jump_gas_cost));
    P cost_filter = lv[OPCODE_BITS] + lv[OPCODE_BITS + 4] - lv[OPCODE_BITS] * lv[OPCODE_BITS + 4];
    P binary_op_cost = P::from_u64(G_LOW) + cost_filter * (P::from_u64(G_VERYLOW) - P::from_u64(G_LOW));
    yc.constraint_transition(lv[OP_BINARY_OP] * (gas_diff - binary_op_cost));
    P ternary_op_cost = P::from_u64(G_MID) - lv[OPCODE_BITS + 1] * P::from_u64(G_MID);
    yc.constraint_transition(lv[OP_TERNARY_OP] * (gas_diff - ternary_op_cost));
    P not_pop_cost = (one - lv[OPCODE_BITS]) * P::from_u64(G_BASE) + lv[OPCODE_BITS] * P::from_u64(G_VERYLOW);
    yc.constraint_transition(lv[OP_NOT_POP] * (gas_diff - not_pop_cost));
    P jdkg_cost = lv[OPCODE_BITS + 1] * P::from_u64(G_JUMPDEST) + (one - lv[OPCODE_BITS + 1]) * P::from_u64(KERNEL_ONLY);
    yc.constraint_transition(lv[OP_JUMPDEST_KECCAK_GENERAL] * (gas_diff - jdkg_cost));
    P ppi_cost = (one - lv[OPCODE_BITS + 7]) * P::from_u64(G_VERYLOW) + lv[OPCODE_BITS + 7] * P::from_u64(KERNEL_ONLY);
    yc.constraint_transition(lv[OP_PUSH_PROVER_INPUT] * (gas_diff - ppi_cost));
    // eval_packed_init
    P is_cpu_cycle = P::zero(), is_cpu_cycle_next = P::zero();
    for (uint32_t op = OP_BEGIN; op < OP_END; op++) { is_cpu_cycle = is_cpu_cycle + lv[op]; is_cpu_cycle_next = is_cpu_cycle_next + nv[op]; }
    P init_filter = (is_cpu_cycle - one) * is_cpu_cycle_next;
    yc.constraint_transition(init_filter * nv[GAS]);
}

// ---- halt.rs:16-52 ------------------------------------------------------------------------------------------------------
template <class P, class V, class CC>
ZKS_HD void eval_halt(const V& lv, const V& nv, CC& yc, const TableParams& prm) {
    const P one = P::one();
    P is_cpu_cycle = P::zero(), is_cpu_cycle_next = P::zero();
    for (uint32_t op = OP_BEGIN; op < OP_END; op++) { is_cpu_cycle = is_cpu_cycle + lv[op]; is_cpu_cycle_next = is_cpu_cycle_next + nv[op]; }
    P halt_state = one - is_cpu_cycle, next_halt_state = one - is_cpu_cycle_next;
    yc.constraint(halt_state * (halt_state - one));
    yc.constraint_transition(halt_state * (next_halt_state - one));
    yc.constraint(halt_state * (lv[IS_KERNEL_MODE] - one));
    for (uint32_t i = 0; i < NUM_GP_CHANNELS; i++) yc.constraint(halt_state * lv[ch_used(i)]);
    yc.constraint_last_row(halt_state - one);
    yc.constraint(halt_state * (lv[PROGRAM_COUNTER] - P::from_u64(prm.halt_final)));
}

// ---- jumps.rs ------------------------------------------------------------------------------------------------------------
template <class P, class V, class CC>
ZKS_HD void eval_jumps(const V& lv, const V& nv, CC& yc) {
    const P one = P::one();
    {   // EXIT_KERNEL
        P filter = lv[OP_EXIT_KERNEL];
        yc.constraint_transition(filter * (lv[ch_value(0, 0)] - nv[PROGRAM_COUNTER]));
        yc.constraint_transition(filter * (lv[ch_value(0, 1)] - nv[IS_KERNEL_MODE]));
        yc.constraint_transition(filter * (lv[ch_value(0, 6)] - nv[GAS]));
        yc.constraint(filter * lv[ch_value(0, 7)]);
    }
    // JUMP / JUMPI: dst = channel 0 value, cond = channel 1 value
    P filter = lv[OP_JUMPS];
    P is_jump = filter * (one - lv[OPCODE_BITS]);
    P is_jumpi = filter * lv[OPCODE_BITS];
    P should_jump = lv[G_JUMPS_SHOULD_JUMP], cond_sum_pinv = lv[G_JUMPS_COND_SUM_PINV];
    P len_diff = lv[STACK_LEN] - one - lv[OPCODE_BITS];
    P new_filter = len_diff * filter;
    yc.constraint_transition(new_filter * (nv[ch_used(0)] - one));
    yc.constraint_transition(new_filter * (nv[ch_is_read(0)] - one));
    yc.constraint_transition(new_filter * (nv[ch_addr_context(0)] - nv[CONTEXT]));
    yc.constraint_transition(new_filter * (nv[ch_addr_segment(0)] - P::from_u64(SEG_STACK)));
    P addr_virtual = nv[STACK_LEN] - one;
    yc.constraint_transition(new_filter * (nv[ch_addr_virtual(0)] - addr_virtual));
    yc.constraint(filter * (len_diff * lv[G_STACK_INV] - lv[G_STACK_INV_AUX]));
    P empty_stack_filter = filter * (lv[G_STACK_INV_AUX] - one);
    yc.constraint_transition(empty_stack_filter * nv[ch_used(0)]);
    yc.constraint(is_jump * (lv[ch_value(1, 0)] - one));
    for (uint32_t i = 1; i < VALUE_LIMBS; i++) yc.constraint(is_jump * lv[ch_value(1, i)]);
    yc.constraint(filter * should_jump * (should_jump - one));
    P cond_sum = P::zero();
    for (uint32_t i = 0; i < VALUE_LIMBS; i++) cond_sum = cond_sum + lv[ch_value(1, i)];
    yc.constraint(filter * (should_jump - one) * cond_sum);
    yc.constraint(filter * (cond_sum_pinv * cond_sum - should_jump));
    P dst_hi_sum = P::zero();
    for (uint32_t i = 1; i < VALUE_LIMBS; i++) dst_hi_sum = dst_hi_sum + lv[ch_value(0, i)];
    yc.constraint(filter * should_jump * dst_hi_sum);
    {   // JUMPDEST flag channel = mem_channels[NUM_GP_CHANNELS - 1]
        const uint32_t jc = NUM_GP_CHANNELS - 1;
        yc.constraint(filter * (lv[ch_value(jc, 0)] - one));
        yc.constraint(filter * (lv[ch_used(jc)] - should_jump * (one - lv[IS_KERNEL_MODE])));
        yc.constraint(filter * (lv[ch_is_read(jc)] - one));
        yc.constraint(filter * (lv[ch_addr_context(jc)] - lv[CONTEXT]));
        yc.constraint(filter * (lv[ch_addr_segment(jc)] - P::from_u64(SEG_JUMPDEST_BITS)));
        yc.constraint(filter * (lv[ch_addr_virtual(jc)] - lv[ch_value(0, 0)]));
    }
    // mem_channels[2..NUM_GP_CHANNELS-1] is empty
    yc.constraint(filter * lv[PC_USED]);
    yc.constraint(is_jump * lv[ch_used(1)]);
    yc.constraint_transition(is_jump * (nv[STACK_LEN] - lv[STACK_LEN] + one));
    yc.constraint_transition(is_jumpi * (nv[STACK_LEN] - lv[STACK_LEN] + P::from_u64(2)));
    P fallthrough_dst = lv[PROGRAM_COUNTER] + one;
    P jump_dest = lv[ch_value(0, 0)];
    yc.constraint_transition(filter * (should_jump - one) * (nv[PROGRAM_COUNTER] - fallthrough_dst));
    yc.constraint_transition(filter * should_jump * (nv[PROGRAM_COUNTER] - jump_dest));
}

// ---- membus.rs:42-58 ------------------------------------------------------------------------------------------------------
template <class P, class V, class CC>
ZKS_HD void eval_membus(const V& lv, CC& yc) {
    const P one = P::one();
    yc.constraint(lv[CODE_CONTEXT] - (one - lv[IS_KERNEL_MODE]) * lv[CONTEXT]);
    for (uint32_t i = 0; i < NUM_GP_CHANNELS; i++) yc.constraint(lv[ch_used(i)] * (lv[ch_used(i)] - one));
    yc.constraint(lv[PC_USED] * (lv[PC_USED] - one));
}

// ---- memio.rs ---------------------------------------------------------------------------------------------------------------
template <class P, class V, class CC>
ZKS_HD void eval_memio(const V& lv, const V& nv, CC& yc) {
    const P one = P::one();
    {   // MLOAD_GENERAL: address in channel 0 value (virt, segment, ctx), load through channel 1
        P filter = lv[OP_M_OP_GENERAL] * lv[OPCODE_BITS];
        yc.constraint(filter * (lv[ch_used(1)] - one));
        yc.constraint(filter * (lv[ch_is_read(1)] - one));
        yc.constraint(filter * (lv[ch_addr_context(1)] - lv[ch_value(0, 2)]));
        yc.constraint(filter * (lv[ch_addr_segment(1)] - lv[ch_value(0, 1)]));
        yc.constraint(filter * (lv[ch_addr_virtual(1)] - lv[ch_value(0, 0)]));
        for (uint32_t i = 0; i < VALUE_LIMBS; i++) yc.constraint(filter * (lv[ch_value(1, i)] - nv[ch_value(0, i)]));
        for (uint32_t c = 2; c < NUM_GP_CHANNELS; c++) yc.constraint(filter * lv[ch_used(c)]);
        yc.constraint(filter * lv[PC_USED]);
        eval_stack_one<P>(lv, nv, filter, StackBehavior{1, true, false}, yc);   // MLOAD_GENERAL_OP
    }
    {   // MSTORE_GENERAL: address in channel 1 value, store through the partial channel
        P filter = lv[OP_M_OP_GENERAL] * (lv[OPCODE_BITS] - one);
        yc.constraint(filter * (lv[PC_USED] - one));
        yc.constraint(filter * lv[PC_IS_READ]);
        yc.constraint(filter * (lv[PC_ADDR_CONTEXT] - lv[ch_value(1, 2)]));
        yc.constraint(filter * (lv[PC_ADDR_SEGMENT] - lv[ch_value(1, 1)]));
        yc.constraint(filter * (lv[PC_ADDR_VIRTUAL] - lv[ch_value(1, 0)]));
        for (uint32_t c = 2; c < NUM_GP_CHANNELS; c++) yc.constraint(filter * lv[ch_used(c)]);
        {   // pops: for i in 1..2
            const uint32_t i = 1;
            yc.constraint(filter * (lv[ch_used(i)] - one));
            yc.constraint(filter * (lv[ch_is_read(i)] - one));
            yc.constraint(filter * (lv[ch_addr_context(i)] - lv[CONTEXT]));
            yc.constraint(filter * (lv[ch_addr_segment(i)] - P::from_u64(SEG_STACK)));
            P addr_virtual = lv[STACK_LEN] - P::from_u64(i + 1);
            yc.constraint(filter * (lv[ch_addr_virtual(i)] - addr_virtual));
        }
        P len_diff = lv[STACK_LEN] - P::from_u64(2);
        yc.constraint(lv[OP_M_OP_GENERAL] * (len_diff * lv[G_STACK_INV] - lv[G_STACK_INV_AUX]));
        P is_top_read = lv[G_STACK_INV_AUX] * (one - lv[OPCODE_BITS]);
        yc.constraint(lv[OP_M_OP_GENERAL] * (lv[G_STACK_INV_AUX_2] - is_top_read));
        P new_filter = lv[OP_M_OP_GENERAL] * lv[G_STACK_INV_AUX_2];
        yc.constraint_transition(new_filter * (nv[ch_used(0)] - one));
        yc.constraint_transition(new_filter * (nv[ch_is_read(0)] - one));
        yc.constraint_transition(new_filter * (nv[ch_addr_context(0)] - nv[CONTEXT]));
        yc.constraint_transition(new_filter * (nv[ch_addr_segment(0)] - P::from_u64(SEG_STACK)));
        P addr_virtual = nv[STACK_LEN] - one;
        yc.constraint_transition(new_filter * (nv[ch_addr_virtual(0)] - addr_virtual));
        yc.constraint(lv[OP_M_OP_GENERAL] * (lv[G_STACK_INV_AUX] - one) * nv[ch_used(0)]);
        yc.constraint(lv[OP_M_OP_GENERAL] * lv[OPCODE_BITS] * nv[ch_used(0)]);
    }
}

// ---- modfp254.rs:19-34, pc.rs:10-22, push0.rs:10-20 -----------------------------------------------------------------------
template <class P, class V, class CC>
ZKS_HD void eval_modfp254(const V& lv, CC& yc) {
    const uint64_t P_LIMBS[8] = {0xd87cfd47, 0x3c208c16, 0x6871ca8d, 0x97816a91, 0x8181585d, 0xb85045b6, 0xe131a029, 0x30644e72};
    P filter = lv[OP_FP254_OP];
    for (uint32_t i = 0; i < VALUE_LIMBS; i++) yc.constraint(filter * (lv[ch_value(2, i)] - P::from_u64(P_LIMBS[i])));
}
template <class P, class V, class CC>
ZKS_HD void eval_pc(const V& lv, const V& nv, CC& yc) {
    P filter = lv[OP_PC_PUSH0] * (P::one() - lv[OPCODE_BITS]);
    yc.constraint(filter * (nv[ch_value(0, 0)] - lv[PROGRAM_COUNTER]));
    for (uint32_t i = 1; i < VALUE_LIMBS; i++) yc.constraint(filter * nv[ch_value(0, i)]);
}
template <class P, class V, class CC>
ZKS_HD void eval_push0(const V& lv, const V& nv, CC& yc) {
    P filter = lv[OP_PC_PUSH0] * lv[OPCODE_BITS];
    for (uint32_t i = 0; i < VALUE_LIMBS; i++) yc.constraint(filter * nv[ch_value(0, i)]);
}

// ---- shift.rs:14-63 ---------------------------------------------------------------------------------------------------------
template <class P, class V, class CC>
ZKS_HD void eval_shift(const V& lv, CC& yc) {
    const P one = P::one();
    P is_shift = lv[OP_SHIFT];
    // displacement = channel 0, two_exp = channel 2
    P high_limbs_are_zero = lv[ch_used(2)];
    yc.constraint(is_shift * high_limbs_are_zero * (lv[ch_is_read(2)] - one));
    P high_limbs_sum = P::zero();
    for (uint32_t i = 1; i < VALUE_LIMBS; i++) high_limbs_sum = high_limbs_sum + lv[ch_value(0, i)];
    P t = high_limbs_sum * lv[G_SHIFT_HIGH_LIMB_SUM_INV] - (one - high_limbs_are_zero);
    yc.constraint(is_shift * t);
    yc.constraint(is_shift * high_limbs_sum * high_limbs_are_zero);
    yc.constraint(is_shift * lv[ch_addr_context(2)]);
    yc.constraint(is_shift * (lv[ch_addr_segment(2)] - P::from_u64(13)));   // Segment::ShiftTable
    yc.constraint(is_shift * (lv[ch_addr_virtual(2)] - lv[ch_value(0, 0)]));
    // mem_channels[3..NUM_GP_CHANNELS] is empty
}

// ---- simple_logic/{not.rs:15-32, eq_iszero.rs:46-102} -----------------------------------------------------------------------
template <class P, class V, class CC>
ZKS_HD void eval_simple_logic(const V& lv, const V& nv, CC& yc) {
    const P one = P::one();
    {   // NOT
        P filter = lv[OP_NOT_POP] * lv[OPCODE_BITS];
        for (uint32_t i = 0; i < VALUE_LIMBS; i++)
            yc.constraint(filter * (nv[ch_value(0, i)] + lv[ch_value(0, i)] - P::from_u64(0xFFFFFFFFULL)));
        eval_stack_one<P>(lv, nv, filter, StackBehavior{1, true, true}, yc);   // BASIC_UNARY_OP
    }
    {   // EQ / ISZERO
        P eq_filter = lv[OP_EQ_ISZERO] * (one - lv[OPCODE_BITS]);
        P iszero_filter = lv[OP_EQ_ISZERO] * lv[OPCODE_BITS];
        P f = lv[OP_EQ_ISZERO];
        P equal = nv[ch_value(0, 0)];
        P unequal = one - equal;
        yc.constraint(f * equal * unequal);
        for (uint32_t i = 1; i < VALUE_LIMBS; i++) yc.constraint(f * nv[ch_value(0, i)]);
        for (uint32_t i = 0; i < VALUE_LIMBS; i++) yc.constraint(iszero_filter * lv[ch_value(1, i)]);
        for (uint32_t i = 0; i < VALUE_LIMBS; i++) yc.constraint(f * equal * (lv[ch_value(0, i)] - lv[ch_value(1, i)]));
        P dot = P::zero();
        for (uint32_t i = 0; i < VALUE_LIMBS; i++) dot = dot + (lv[ch_value(0, i)] - lv[ch_value(1, i)]) * lv[G_LOGIC_DIFF_PINV + i];
        yc.constraint(f * (dot - unequal));
        eval_stack_one<P>(lv, nv, eq_filter, StackBehavior{2, true, true}, yc);       // EQ_STACK_BEHAVIOR
        eval_stack_one<P>(lv, nv, iszero_filter, StackBehavior{1, true, true}, yc);   // IS_ZERO_STACK_BEHAVIOR
    }
}

// ---- stack.rs:316-412 eval_packed --------------------------------------------------------------------------------------------
template <class P, class V, class CC>
ZKS_HD void eval_stack(const V& lv, const V& nv, CC& yc) {
    const P one = P::one();
    // STACK_BEHAVIORS / MIGHT_OVERFLOW in OpsColumnsView field order (stack.rs:23-180); num_pops < 0 = None
    const int SB_POPS[NUM_OPS] = {2, 3, 2, -1, 2, -1, 2, -1, -1, 0, -1, -1, 2, 1, -1, 0, 0, 0};
    const bool SB_PUSH[NUM_OPS] = {true, true, true, false, true, false, true, false, false, true, false, false, true, false, false, true, true, true};
    const bool SB_DIS[NUM_OPS] = {true, true, true, false, true, false, false, false, false, true, false, false, false, true, false, true, false, false};
    const bool OVERFLOW[NUM_OPS] = {false, false, false, false, false, false, false, false, false, true, true, false, false, true, false, true, false, false};
    for (uint32_t i = 0; i < NUM_OPS; i++) {
        P op = lv[OP_BEGIN + i];
        if (SB_POPS[i] >= 0) eval_stack_one<P>(lv, nv, op, StackBehavior{(uint32_t)SB_POPS[i], SB_PUSH[i], SB_DIS[i]}, yc);
        if (OVERFLOW[i]) {
            P diff = nv[STACK_LEN] - P::from_u64(1025);
            P lhs = diff * lv[G_STACK_LEN_BOUNDS_AUX];
            P rhs = one - nv[IS_KERNEL_MODE];
            yc.constraint_transition(op * (lhs - rhs));
        }
    }
    P jumpdest_filter = lv[OP_JUMPDEST_KECCAK_GENERAL] * lv[OPCODE_BITS + 1];
    eval_stack_one<P>(lv, nv, jumpdest_filter, StackBehavior{0, false, true}, yc);          // JUMPDEST_OP
    P keccak_general_filter = lv[OP_JUMPDEST_KECCAK_GENERAL] * (one - lv[OPCODE_BITS + 1]);
    eval_stack_one<P>(lv, nv, keccak_general_filter, StackBehavior{2, true, true}, yc);     // KECCAK_GENERAL_OP
    // POP
    P len_diff = lv[STACK_LEN] - one;
    yc.constraint(lv[OP_NOT_POP] * (len_diff * lv[G_STACK_INV] - lv[G_STACK_INV_AUX]));
    P is_top_read = lv[G_STACK_INV_AUX] * (one - lv[OPCODE_BITS]);
    yc.constraint(lv[OP_NOT_POP] * (lv[G_STACK_INV_AUX_2] - is_top_read));
    P new_filter = lv[OP_NOT_POP] * lv[G_STACK_INV_AUX_2];
    yc.constraint_transition(new_filter * (nv[ch_used(0)] - one));
    yc.constraint_transition(new_filter * (nv[ch_is_read(0)] - one));
    yc.constraint_transition(new_filter * (nv[ch_addr_context(0)] - nv[CONTEXT]));
    yc.constraint_transition(new_filter * (nv[ch_addr_segment(0)] - P::from_u64(SEG_STACK)));
    P addr_virtual = nv[STACK_LEN] - one;
    yc.constraint_transition(new_filter * (nv[ch_addr_virtual(0)] - addr_virtual));
    yc.constraint(lv[OP_NOT_POP] * (lv[G_STACK_INV_AUX_2] - one) * nv[ch_used(0)]);
    for (uint32_t c = 1; c < NUM_GP_CHANNELS; c++) yc.constraint(lv[OP_NOT_POP] * (lv[OPCODE_BITS] - one) * lv[ch_used(c)]);
    yc.constraint(lv[OP_NOT_POP] * (lv[OPCODE_BITS] - one) * lv[PC_USED]);
    yc.constraint_transition(lv[OP_NOT_POP] * (lv[OPCODE_BITS] - one) * (nv[STACK_LEN] - lv[STACK_LEN] + one));
}

// ---- syscalls_exceptions.rs:23-134 ---------------------------------------------------------------------------------------------
template <class P, class V, class CC>
ZKS_HD void eval_syscalls_exceptions(const V& lv, const V& nv, CC& yc, const TableParams& prm) {
    const P one = P::one();
    const uint64_t BYTES_PER_OFFSET = 3, EXC_STOP_CODE = 6;
    P filter_syscall = lv[OP_SYSCALL], filter_exception = lv[OP_EXCEPTION];
    P total_filter = filter_syscall + filter_exception;
    yc.constraint(filter_syscall * (filter_syscall - one));
    yc.constraint(filter_exception * (filter_exception - one));
    P exc_code = P::zero();
    for (uint32_t i = 0; i < 3; i++) exc_code = exc_code + lv[G_EXC_CODE_BITS + i] * P::from_u64(1ULL << i);
    P exc_stop_code = P::from_u64(EXC_STOP_CODE);
    yc.constraint(filter_exception * (exc_code - exc_stop_code) * lv[IS_KERNEL_MODE]);
    for (uint32_t i = 0; i < 3; i++) { P bit = lv[G_EXC_CODE_BITS + i]; yc.constraint(filter_exception * bit * (bit - one)); }
    P opcode = P::zero();
    for (uint32_t i = 0; i < 8; i++) opcode = opcode + lv[OPCODE_BITS + i] * P::from_u64(1ULL << i);
    P opcode_handler_addr_start = P::from_u64(prm.syscall_jumptable) + opcode * P::from_u64(BYTES_PER_OFFSET);
    P exc_handler_addr_start = P::from_u64(prm.exception_jumptable) + exc_code * P::from_u64(BYTES_PER_OFFSET);
    // jumpdest channel = channel 1
    yc.constraint(total_filter * lv[ch_used(1)]);
    yc.constraint(total_filter * (lv[ch_is_read(1)] - one));
    yc.constraint(total_filter * lv[ch_addr_context(1)]);
    yc.constraint(total_filter * (lv[ch_addr_segment(1)] - P::from_u64(SEG_CODE)));
    yc.constraint(filter_syscall * (lv[ch_addr_virtual(1)] - opcode_handler_addr_start));
    yc.constraint(filter_exception * (lv[ch_addr_virtual(1)] - exc_handler_addr_start));
    for (uint32_t i = 1; i < VALUE_LIMBS; i++) yc.constraint(total_filter * lv[ch_value(1, i)]);
    for (uint32_t c = 2; c < NUM_GP_CHANNELS; c++) yc.constraint(total_filter * lv[ch_used(c)]);
    yc.constraint_transition(total_filter * (nv[PROGRAM_COUNTER] - lv[ch_value(1, 0)]));
    yc.constraint_transition(total_filter * (nv[IS_KERNEL_MODE] - one));
    yc.constraint_transition(total_filter * nv[GAS]);
    yc.constraint(filter_syscall * (nv[ch_value(0, 0)] - (lv[PROGRAM_COUNTER] + one)));
    yc.constraint(filter_exception * (nv[ch_value(0, 0)] - lv[PROGRAM_COUNTER]));
    yc.constraint(filter_syscall * (nv[ch_value(0, 1)] - lv[IS_KERNEL_MODE]));
    yc.constraint(total_filter * (nv[ch_value(0, 6)] - lv[GAS]));
    yc.constraint(total_filter * nv[ch_value(0, 7)]);
    yc.constraint(filter_exception * (exc_code - exc_stop_code) * nv[ch_value(0, 1)]);
    for (uint32_t i = 2; i < 6; i++) yc.constraint(total_filter * nv[ch_value(0, i)]);
}

// ---- cpu_stark.rs:594-626: the dispatcher -----------------------------------------------------------------------------------------
template <class P, class V, class CC>
ZKS_HD void eval(const V& lv, const V& nv, CC& yc, const TableParams& prm) {
    eval_byte_unpacking<P>(lv, nv, yc);
    ZKS_SYNC();
    eval_clock<P>(lv, nv, yc);
    ZKS_SYNC();
    eval_contextops<P>(lv, nv, yc);
    ZKS_SYNC();
    eval_control_flow<P>(lv, nv, yc, prm);
    ZKS_SYNC();
    eval_decode<P>(lv, yc);
    ZKS_SYNC();
    eval_dup_swap<P>(lv, nv, yc);
    ZKS_SYNC();
    eval_gas<P>(lv, nv, yc);
    ZKS_SYNC();
    eval_halt<P>(lv, nv, yc, prm);
    ZKS_SYNC();
    eval_jumps<P>(lv, nv, yc);
    ZKS_SYNC();
    eval_membus<P>(lv, yc);
    ZKS_SYNC();
    eval_memio<P>(lv, nv, yc);
    ZKS_SYNC();
    eval_modfp254<P>(lv, yc);
    ZKS_SYNC();
    eval_pc<P>(lv, nv, yc);
    ZKS_SYNC();
    eval_push0<P>(lv, nv, yc);
    ZKS_SYNC();
    eval_shift<P>(lv, yc);
    ZKS_SYNC();
    eval_simple_logic<P>(lv, nv, yc);
    ZKS_SYNC();
    eval_stack<P>(lv, nv, yc);
    ZKS_SYNC();
    eval_syscalls_exceptions<P>(lv, nv, yc, prm);
}

// ---- cross-table lookups (cpu_stark.rs:33-470) --------------------------------------------------------------------------------------
inline std::vector<uint32_t> value_cols(uint32_t ch) { std::vector<uint32_t> v; for (uint32_t i = 0; i < VALUE_LIMBS; i++) v.push_back(ch_value(ch, i)); return v; }
inline void extend(std::vector<Column>& a, const std::vector<Column>& b) { a.insert(a.end(), b.begin(), b.end()); }
inline std::vector<Column> singles_of(const std::vector<uint32_t>& cs) { std::vector<Column> v; for (uint32_t c : cs) v.push_back(Column::single(c)); return v; }
inline std::vector<Column> singles_next_row_of(const std::vector<uint32_t>& cs) { std::vector<Column> v; for (uint32_t c : cs) v.push_back(Column::single_next_row(c)); return v; }
inline std::vector<uint32_t> opcode_bit_cols(uint32_t lo, uint32_t hi) { std::vector<uint32_t> v; for (uint32_t i = lo; i < hi; i++) v.push_back(OPCODE_BITS + i); return v; }
// timestamp = (clock - 1) * NUM_CHANNELS + 1
inline Column ctl_timestamp() { return Column::linear_combination_with_constant({{CLOCK, NUM_CHANNELS}}, canon(GL_MOD + 1 - NUM_CHANNELS)); }
// mem_time_and_channel: clock * NUM_CHANNELS + channel - NUM_CHANNELS + 1
inline Column mem_time_and_channel(uint32_t channel) {
    return Column::linear_combination_with_constant({{CLOCK, NUM_CHANNELS}}, canon(GL_MOD + channel + 1 - NUM_CHANNELS));
}

inline std::vector<Column> ctl_data_keccak_sponge() {
    // get_addr(channel 0): context = value[2], segment = value[1], virt = value[0]
    std::vector<Column> cols = {Column::single(ch_value(0, 2)), Column::single(ch_value(0, 1)), Column::single(ch_value(0, 0)),
                                Column::single(ch_value(1, 0)), ctl_timestamp()};
    extend(cols, singles_next_row_of(value_cols(0)));
    return cols;
}
inline Filter ctl_filter_keccak_sponge() {
    return Filter({{Column::single(OP_JUMPDEST_KECCAK_GENERAL), Column::linear_combination_with_constant({{OPCODE_BITS + 1, GL_MOD - 1}}, 1)}}, {});
}
inline std::vector<Column> ctl_data_binops() {
    std::vector<Column> res = singles_of(value_cols(0));
    extend(res, singles_of(value_cols(1)));
    extend(res, singles_next_row_of(value_cols(0)));
    return res;
}
inline std::vector<Column> ctl_data_ternops() {
    std::vector<Column> res = singles_of(value_cols(0));
    extend(res, singles_of(value_cols(1)));
    extend(res, singles_of(value_cols(2)));
    extend(res, singles_next_row_of(value_cols(0)));
    return res;
}
inline std::vector<Column> ctl_data_logic() {
    std::vector<Column> res = {Column::le_bits(opcode_bit_cols(0, 8))};
    extend(res, ctl_data_binops());
    return res;
}
inline Filter ctl_filter_logic() { return Filter::new_simple(Column::single(OP_LOGIC_OP)); }
inline TableWithColumns ctl_arithmetic_base_rows() {
    std::vector<Column> columns = {Column::le_bits(opcode_bit_cols(0, 8))};
    extend(columns, ctl_data_ternops());
    return TableWithColumns(2, columns,
                            Filter({{Column::single(OP_PUSH_PROVER_INPUT), Column::single(OPCODE_BITS + 7)}},
                                   {Column::sum({OP_BINARY_OP, OP_FP254_OP, OP_TERNARY_OP, OP_SHIFT, OP_SYSCALL, OP_EXCEPTION})}));
}
inline TableWithColumns ctl_context_pruning_looked() {
    return TableWithColumns(2, {Column::single(CONTEXT)}, Filter({{Column::single(OP_CONTEXT_OP), Column::single(G_PRUNING_FLAG)}}, {}));
}
inline std::vector<Column> ctl_data_byte_packing() {
    std::vector<Column> res = {Column::constant_col(1)};   // is_read
    extend(res, ctl_data_keccak_sponge());
    return res;
}
inline Filter ctl_filter_byte_packing() { return Filter({{Column::single(OP_M_OP_32BYTES), Column::single(OPCODE_BITS + 5)}}, {}); }
inline std::vector<Column> ctl_data_byte_unpacking() {
    std::vector<Column> res = {Column::constant_col(0), Column::single(ch_value(0, 2)), Column::single(ch_value(0, 1)),
                               Column::single(ch_value(0, 0))};
    // len = new_offset - virt
    res.push_back(Column::linear_combination_and_next_row_with_constant({{ch_value(0, 0), GL_MOD - 1}}, {{ch_value(0, 0), 1}}, 0));
    res.push_back(ctl_timestamp());
    extend(res, singles_of(value_cols(1)));
    return res;
}
inline Filter ctl_filter_byte_unpacking() {
    return Filter({{Column::single(OP_M_OP_32BYTES), Column::linear_combination_with_constant({{OPCODE_BITS + 5, GL_MOD - 1}}, 1)}}, {});
}
inline std::vector<Column> ctl_data_jumptable_read() {
    std::vector<Column> res = {Column::constant_col(1), Column::single(ch_addr_context(1)), Column::single(ch_addr_segment(1)),
                               Column::single(ch_addr_virtual(1)), Column::constant_col(3), ctl_timestamp()};
    extend(res, singles_of(value_cols(1)));
    return res;
}
inline Filter ctl_filter_syscall_exceptions() { return Filter::new_simple(Column::sum({OP_SYSCALL, OP_EXCEPTION})); }
inline std::vector<Column> ctl_data_byte_packing_push() {
    std::vector<Column> res = {Column::constant_col(1), Column::single(CODE_CONTEXT), Column::constant_col(0 /* Segment::Code as usize */),
                               Column::linear_combination_with_constant({{PROGRAM_COUNTER, 1}}, 1),
                               Column::le_bits_with_constant(opcode_bit_cols(0, 5), 1), ctl_timestamp()};
    extend(res, singles_next_row_of(value_cols(0)));
    return res;
}
inline Filter ctl_filter_byte_packing_push() { return Filter({{Column::single(G_PUSH_IS_NOT_KERNEL), Column::single(OP_PUSH_PROVER_INPUT)}}, {}); }

static const uint32_t MEM_CODE_CHANNEL_IDX = 0, MEM_GP_CHANNELS_IDX_START = 1;
inline std::vector<Column> ctl_data_code_memory() {
    std::vector<Column> cols = {Column::constant_col(1), Column::single(CODE_CONTEXT), Column::constant_col(SEG_CODE), Column::single(PROGRAM_COUNTER),
                                Column::le_bits(opcode_bit_cols(0, 8))};
    for (uint32_t i = 0; i < VALUE_LIMBS - 1; i++) cols.push_back(Column::constant_col(0));
    cols.push_back(mem_time_and_channel(MEM_CODE_CHANNEL_IDX));
    return cols;
}
inline std::vector<Column> ctl_data_gp_memory(uint32_t ch) {
    std::vector<Column> cols = {Column::single(ch_is_read(ch)), Column::single(ch_addr_context(ch)), Column::single(ch_addr_segment(ch)),
                                Column::single(ch_addr_virtual(ch))};
    extend(cols, singles_of(value_cols(ch)));
    cols.push_back(mem_time_and_channel(MEM_GP_CHANNELS_IDX_START + ch));
    return cols;
}
inline std::vector<Column> ctl_data_partial_memory() {
    std::vector<Column> cols = {Column::single(PC_IS_READ), Column::single(PC_ADDR_CONTEXT), Column::single(PC_ADDR_SEGMENT),
                                Column::single(PC_ADDR_VIRTUAL)};
    extend(cols, singles_of(value_cols(0)));
    cols.push_back(mem_time_and_channel(MEM_GP_CHANNELS_IDX_START + NUM_GP_CHANNELS));
    return cols;
}
static const uint64_t CONTEXT_METADATA_STACK_SIZE = 11;   // ContextMetadata::StackSize.unscale()
inline std::vector<Column> ctl_data_memory_old_sp_write_set_context() {
    std::vector<Column> cols = {Column::constant_col(0), Column::single(CONTEXT), Column::constant_col(SEG_CONTEXT_METADATA),
                                Column::constant_col(CONTEXT_METADATA_STACK_SIZE),
                                Column::linear_combination_with_constant({{STACK_LEN, 1}}, GL_MOD - 1)};
    for (uint32_t i = 0; i < VALUE_LIMBS - 1; i++) cols.push_back(Column::constant_col(0));
    cols.push_back(mem_time_and_channel(MEM_GP_CHANNELS_IDX_START + 1));
    return cols;
}
inline std::vector<Column> ctl_data_memory_new_sp_read_set_context() {
    std::vector<Column> cols = {Column::constant_col(1), Column::single(ch_value(0, 2)), Column::constant_col(SEG_CONTEXT_METADATA),
                                Column::constant_col(CONTEXT_METADATA_STACK_SIZE), Column::single_next_row(STACK_LEN)};
    for (uint32_t i = 0; i < VALUE_LIMBS - 1; i++) cols.push_back(Column::constant_col(0));
    cols.push_back(mem_time_and_channel(MEM_GP_CHANNELS_IDX_START + 2));
    return cols;
}
inline Filter ctl_filter_code_memory() {
    std::vector<uint32_t> ops; for (uint32_t op = OP_BEGIN; op < OP_END; op++) ops.push_back(op);
    return Filter::new_simple(Column::sum(ops));
}
inline Filter ctl_filter_gp_memory(uint32_t ch) { return Filter::new_simple(Column::single(ch_used(ch))); }
inline Filter ctl_filter_partial_memory() { return Filter::new_simple(Column::single(PC_USED)); }
inline Filter ctl_filter_set_context() { return Filter({{Column::single(OP_CONTEXT_OP), Column::single(OPCODE_BITS)}}, {}); }
inline std::vector<Lookup> lookups() { return {}; }

}}  // namespace zkstark::cpu
