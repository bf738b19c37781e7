// AllStark registry: the nine tables, their widths, constraint evaluators, lookups and the ten cross-table lookups.
// Source: /root/reference/evm_arithmetization/src/all_stark.rs:34-172 (AllStark, Table, all_cross_table_lookups) and :176-417.
#pragma once
#include "hd.h"
#include "lookup.h"
#include "consumer.h"
#include "table_memcont.h"
#include "table_memory.h"
#include "table_logic.h"
#if __has_include("table_cpu.h") && __has_include("table_byte_packing.h") && __has_include("table_arithmetic.h") && \
    __has_include("table_keccak.h") && __has_include("table_keccak_sponge.h")
#include "table_cpu.h"
#include "table_byte_packing.h"
#include "table_arithmetic.h"
#include "table_keccak.h"
#include "table_keccak_sponge.h"
#define ZKS_ALL_TABLES 1
#else
#define ZKS_ALL_TABLES 0
#endif

namespace zkstark {

enum Table : uint32_t {
    T_ARITHMETIC = 0, T_BYTE_PACKING = 1, T_CPU = 2, T_KECCAK = 3, T_KECCAK_SPONGE = 4, T_LOGIC = 5, T_MEMORY = 6,
    T_MEM_BEFORE = 7, T_MEM_AFTER = 8, NUM_TABLES = 9
};
static const unsigned CONSTRAINT_DEGREE = 3;          // every table: constraint_degree() == 3
static const unsigned NUM_CTLS = 10;
static const unsigned MEMORY_CTL_IDX = 6;              // all_stark.rs:150

inline const char* table_name(uint32_t t) {
    static const char* n[] = {"arithmetic", "byte_packing", "cpu", "keccak", "keccak_sponge", "logic", "memory",
                              "mem_before", "mem_after"};
    return t < NUM_TABLES ? n[t] : "?";
}
inline bool table_is_optional(uint32_t t) {   // OPTIONAL_TABLE_INDICES, all_stark.rs:110-117
    return t == T_BYTE_PACKING || t == T_KECCAK || t == T_KECCAK_SPONGE || t == T_LOGIC || t == T_MEM_AFTER;
}
inline bool table_supported(uint32_t t) {
#if ZKS_ALL_TABLES
    return t < NUM_TABLES;
#else
    return t == T_LOGIC || t == T_MEMORY || t == T_MEM_BEFORE || t == T_MEM_AFTER;
#endif
}
inline uint32_t table_num_columns(uint32_t t) {
    switch (t) {
#if ZKS_ALL_TABLES
        case T_ARITHMETIC: return arithmetic::NUM_COLUMNS;
        case T_BYTE_PACKING: return byte_packing::NUM_COLUMNS;
        case T_CPU: return cpu::NUM_COLUMNS;
        case T_KECCAK: return keccak::NUM_COLUMNS;
        case T_KECCAK_SPONGE: return keccak_sponge::NUM_COLUMNS;
#endif
        case T_LOGIC: return logic::NUM_COLUMNS;
        case T_MEMORY: return memory::NUM_COLUMNS;
        case T_MEM_BEFORE: case T_MEM_AFTER: return memcont::NUM_COLUMNS;
        default: return 0;
    }
}
inline std::vector<Lookup> table_lookups(uint32_t t) {
    switch (t) {
#if ZKS_ALL_TABLES
        case T_ARITHMETIC: return arithmetic::lookups();
        case T_BYTE_PACKING: return byte_packing::lookups();
        case T_KECCAK_SPONGE: return keccak_sponge::lookups();
#endif
        case T_MEMORY: return memory::lookups();
        default: return {};
    }
}

// host-side dispatch (oracle prover + verifier); the CUDA side instantiates one kernel per table instead
template <class P, class V, class CC>
inline void eval_table(uint32_t t, const V& lv, const V& nv, CC& yc, const TableParams& prm) {
    switch (t) {
#if ZKS_ALL_TABLES
        case T_ARITHMETIC: arithmetic::eval<P>(lv, nv, yc); break;
        case T_BYTE_PACKING: byte_packing::eval<P>(lv, nv, yc); break;
        case T_CPU: cpu::eval<P>(lv, nv, yc, prm); break;
        case T_KECCAK: keccak::eval<P>(lv, nv, yc); break;
        case T_KECCAK_SPONGE: keccak_sponge::eval<P>(lv, nv, yc); break;
#endif
        case T_LOGIC: logic::eval<P>(lv, nv, yc); break;
        case T_MEMORY: memory::eval<P>(lv, nv, yc); break;
        case T_MEM_BEFORE: case T_MEM_AFTER: memcont::eval<P>(lv, nv, yc); break;
        default: break;
    }
}

// ---- cross-table lookups, in the order of all_cross_table_lookups() (all_stark.rs:153-172) -------------------
inline CrossTableLookup ctl_mem_before() {
    return CrossTableLookup({TableWithColumns(T_MEMORY, memory::ctl_looking_mem(), memory::ctl_filter_mem_before())},
                            TableWithColumns(T_MEM_BEFORE, memcont::ctl_data(), memcont::ctl_filter()));
}
inline CrossTableLookup ctl_mem_after() {
    return CrossTableLookup({TableWithColumns(T_MEMORY, memory::ctl_looking_mem(), memory::ctl_filter_mem_after())},
                            TableWithColumns(T_MEM_AFTER, memcont::ctl_data(), memcont::ctl_filter()));
}
#if ZKS_ALL_TABLES
inline CrossTableLookup ctl_arithmetic() {
    return CrossTableLookup({cpu::ctl_arithmetic_base_rows()}, arithmetic::ctl_arithmetic_rows());
}
inline CrossTableLookup ctl_byte_packing() {
    return CrossTableLookup(
        {TableWithColumns(T_CPU, cpu::ctl_data_byte_packing(), cpu::ctl_filter_byte_packing()),
         TableWithColumns(T_CPU, cpu::ctl_data_byte_unpacking(), cpu::ctl_filter_byte_unpacking()),
         TableWithColumns(T_CPU, cpu::ctl_data_byte_packing_push(), cpu::ctl_filter_byte_packing_push()),
         TableWithColumns(T_CPU, cpu::ctl_data_jumptable_read(), cpu::ctl_filter_syscall_exceptions())},
        TableWithColumns(T_BYTE_PACKING, byte_packing::ctl_looked_data(), byte_packing::ctl_looked_filter()));
}
inline CrossTableLookup ctl_keccak_sponge() {
    return CrossTableLookup({TableWithColumns(T_CPU, cpu::ctl_data_keccak_sponge(), cpu::ctl_filter_keccak_sponge())},
                            TableWithColumns(T_KECCAK_SPONGE, keccak_sponge::ctl_looked_data(), keccak_sponge::ctl_looked_filter()));
}
inline CrossTableLookup ctl_keccak_inputs() {
    return CrossTableLookup({TableWithColumns(T_KECCAK_SPONGE, keccak_sponge::ctl_looking_keccak_inputs(),
                                              keccak_sponge::ctl_looking_keccak_filter())},
                            TableWithColumns(T_KECCAK, keccak::ctl_data_inputs(), keccak::ctl_filter_inputs()));
}
inline CrossTableLookup ctl_keccak_outputs() {
    return CrossTableLookup({TableWithColumns(T_KECCAK_SPONGE, keccak_sponge::ctl_looking_keccak_outputs(),
                                              keccak_sponge::ctl_looking_keccak_filter())},
                            TableWithColumns(T_KECCAK, keccak::ctl_data_outputs(), keccak::ctl_filter_outputs()));
}
inline CrossTableLookup ctl_logic() {
    std::vector<TableWithColumns> lookers = {TableWithColumns(T_CPU, cpu::ctl_data_logic(), cpu::ctl_filter_logic())};
    for (uint32_t i = 0; i < keccak_sponge::num_logic_ctls(); i++)
        lookers.push_back(TableWithColumns(T_KECCAK_SPONGE, keccak_sponge::ctl_looking_logic(i),
                                           keccak_sponge::ctl_looking_logic_filter()));
    return CrossTableLookup(lookers, TableWithColumns(T_LOGIC, logic::ctl_data(), logic::ctl_filter()));
}
inline CrossTableLookup ctl_memory() {
    std::vector<TableWithColumns> lookers = {
        TableWithColumns(T_CPU, cpu::ctl_data_code_memory(), cpu::ctl_filter_code_memory()),
        TableWithColumns(T_CPU, cpu::ctl_data_partial_memory(), cpu::ctl_filter_partial_memory()),
        TableWithColumns(T_CPU, cpu::ctl_data_memory_old_sp_write_set_context(), cpu::ctl_filter_set_context()),
        TableWithColumns(T_CPU, cpu::ctl_data_memory_new_sp_read_set_context(), cpu::ctl_filter_set_context())};
    for (uint32_t ch = 0; ch < cpu::NUM_GP_CHANNELS; ch++)
        lookers.push_back(TableWithColumns(T_CPU, cpu::ctl_data_gp_memory(ch), cpu::ctl_filter_gp_memory(ch)));
    for (uint32_t i = 0; i < keccak_sponge::KECCAK_RATE_BYTES; i++)
        lookers.push_back(TableWithColumns(T_KECCAK_SPONGE, keccak_sponge::ctl_looking_memory(i),
                                           keccak_sponge::ctl_looking_memory_filter(i)));
    for (uint32_t i = 0; i < 32; i++)
        lookers.push_back(TableWithColumns(T_BYTE_PACKING, byte_packing::ctl_looking_memory(i),
                                           byte_packing::ctl_looking_memory_filter(i)));
    lookers.push_back(TableWithColumns(T_MEM_BEFORE, memcont::ctl_data_memory(), memcont::ctl_filter()));
    return CrossTableLookup(lookers, TableWithColumns(T_MEMORY, memory::ctl_data(), memory::ctl_filter()));
}
inline CrossTableLookup ctl_context_pruning() {
    return CrossTableLookup({memory::ctl_context_pruning_looking()}, cpu::ctl_context_pruning_looked());
}
inline std::vector<CrossTableLookup> all_cross_table_lookups() {
    return {ctl_arithmetic(), ctl_byte_packing(), ctl_keccak_sponge(), ctl_keccak_inputs(), ctl_keccak_outputs(),
            ctl_logic(),      ctl_memory(),       ctl_mem_before(),    ctl_mem_after(),     ctl_context_pruning()};
}

#else
// reduced registry while the remaining tables are being transcribed: only the CTLs whose tables all exist, plus
// the looked side of the memory CTL
inline std::vector<CrossTableLookup> all_cross_table_lookups() {
    std::vector<CrossTableLookup> v(NUM_CTLS);
    for (auto& c : v) { c.looked_table.table = 0xFFFFFFFFu; }
    v[5] = CrossTableLookup({}, TableWithColumns(T_LOGIC, logic::ctl_data(), logic::ctl_filter()));
    v[MEMORY_CTL_IDX] = CrossTableLookup({TableWithColumns(T_MEM_BEFORE, memcont::ctl_data_memory(), memcont::ctl_filter())},
                                         TableWithColumns(T_MEMORY, memory::ctl_data(), memory::ctl_filter()));
    v[7] = ctl_mem_before();
    v[8] = ctl_mem_after();
    return v;
}
#endif

}  // namespace zkstark
