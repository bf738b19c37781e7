// Host-side description of cross-table lookups and logUp range checks.
//
// Mirrors the builder API of starky 1.0.0 lookup.rs / cross_table_lookup.rs (`Column`, `Filter`, `Lookup`,
// `TableWithColumns`, `CrossTableLookup`) as used by /root/reference/evm_arithmetization/src/all_stark.rs:153-417 and by
// every table's ctl_* / lookups() function, so those can be transcribed one to one.  `Flat` turns the descriptions
// into plain arrays that both the CUDA kernels and the oracle interpret.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <vector>
#include <utility>

namespace zkstark {

static const uint64_t GL_MOD = 0xFFFFFFFF00000001ULL;
inline uint64_t canon(uint64_t x) { return x >= GL_MOD ? x - GL_MOD : x; }
inline uint64_t neg_const(uint64_t x) { x = canon(x); return x ? GL_MOD - x : 0; }
inline uint64_t mulmod_const(uint64_t a, uint64_t b) { return (uint64_t)(((unsigned __int128)a * b) % GL_MOD); }

// A linear combination of cells of the current row and of the next row, plus a constant (starky `Column`).
struct Column {
    std::vector<std::pair<uint32_t, uint64_t>> lin, next;
    uint64_t constant = 0;

    static Column single(uint32_t c) { Column r; r.lin.push_back({c, 1}); return r; }
    static Column single_next_row(uint32_t c) { Column r; r.next.push_back({c, 1}); return r; }
    static Column constant_col(uint64_t k) { Column r; r.constant = canon(k); return r; }
    static Column zero() { return constant_col(0); }
    static Column one() { return constant_col(1); }
    template <class It> static std::vector<Column> singles(It b, It e) {
        std::vector<Column> v; for (; b != e; ++b) v.push_back(single((uint32_t)*b)); return v;
    }
    static std::vector<Column> singles(std::initializer_list<uint32_t> cs) { return singles(cs.begin(), cs.end()); }
    static std::vector<Column> singles_range(uint32_t lo, uint32_t hi) {
        std::vector<Column> v; for (uint32_t c = lo; c < hi; c++) v.push_back(single(c)); return v;
    }
    static Column linear_combination_with_constant(const std::vector<std::pair<uint32_t, uint64_t>>& t, uint64_t k) {
        Column r; r.lin = t; for (auto& p : r.lin) p.second = canon(p.second); r.constant = canon(k); return r;
    }
    static Column linear_combination(const std::vector<std::pair<uint32_t, uint64_t>>& t) {
        return linear_combination_with_constant(t, 0);
    }
    static Column linear_combination_and_next_row_with_constant(const std::vector<std::pair<uint32_t, uint64_t>>& t,
                                                                const std::vector<std::pair<uint32_t, uint64_t>>& nx,
                                                                uint64_t k) {
        Column r = linear_combination_with_constant(t, k); r.next = nx; for (auto& p : r.next) p.second = canon(p.second);
        return r;
    }
    // sum_i c_i * 2^i
    static Column le_bits(const std::vector<uint32_t>& cs) {
        Column r; uint64_t w = 1;
        for (uint32_t c : cs) { r.lin.push_back({c, w}); w = mulmod_const(w, 2); }
        return r;
    }
    // sum_i c_i * 2^i + k
    static Column le_bits_with_constant(const std::vector<uint32_t>& cs, uint64_t k) {
        Column r = le_bits(cs); r.constant = canon(k); return r;
    }
    // sum_i c_i * 256^i
    static Column le_bytes(const std::vector<uint32_t>& cs) {
        Column r; uint64_t w = 1;
        for (uint32_t c : cs) { r.lin.push_back({c, w}); w = mulmod_const(w, 256); }
        return r;
    }
    static Column sum(const std::vector<uint32_t>& cs) {
        Column r; for (uint32_t c : cs) r.lin.push_back({c, 1}); return r;
    }
};

// sum of products of two columns + sum of columns; evaluates to 0/1 (starky `Filter`).  Default: always on.
struct Filter {
    std::vector<std::pair<Column, Column>> products;
    std::vector<Column> constants;
    Filter() { constants.push_back(Column::one()); }
    Filter(const std::vector<std::pair<Column, Column>>& p, const std::vector<Column>& c) : products(p), constants(c) {}
    static Filter new_simple(const Column& c) { return Filter({}, {c}); }
};

struct TableWithColumns {
    uint32_t table;
    std::vector<Column> columns;
    Filter filter;
    TableWithColumns() : table(0) {}
    TableWithColumns(uint32_t t, const std::vector<Column>& c, const Filter& f) : table(t), columns(c), filter(f) {}
};

struct CrossTableLookup {
    std::vector<TableWithColumns> looking_tables;
    TableWithColumns looked_table;
    CrossTableLookup() {}
    CrossTableLookup(const std::vector<TableWithColumns>& l, const TableWithColumns& d) : looking_tables(l), looked_table(d) {}
};

// logUp range check inside one table (starky `Lookup`)
struct Lookup {
    std::vector<Column> columns;
    Column table_column;
    Column frequencies_column;
    std::vector<Filter> filter_columns;
    size_t num_helper_columns(unsigned constraint_degree) const {   // helpers + Z
        size_t d = constraint_degree - 1;
        return (columns.size() + d - 1) / d + 1;
    }
};

// ---------------------------------------------------------------------------------------------------------------
// Flat (POD) form
// ---------------------------------------------------------------------------------------------------------------
struct ColRec { uint32_t lin_begin, lin_end, next_begin, next_end; uint64_t constant; };
struct FilterRec { uint32_t prod_begin, prod_end;     // pairs: prod[2k], prod[2k+1] are ColRec ids
                   uint32_t const_begin, const_end; };   // consts[k] are ColRec ids
// one (columns, filter) pair of a lookup: `ncols` ColRec ids starting at col_begin in `col_ids`
struct EntryRec { uint32_t col_begin, col_end; uint32_t filter; };
// one CtlZData of a table: entries [entry_begin, entry_end), helper columns count, which challenge pair
struct CtlZRec { uint32_t entry_begin, entry_end; uint32_t num_helpers; uint32_t challenge; };
// one in-table Lookup: entries (single-column) [entry_begin, entry_end), table/frequency ColRec ids
struct LookupRec { uint32_t entry_begin, entry_end; uint32_t table_col, freq_col; uint32_t num_helpers; /* without Z */ };

struct Flat {
    std::vector<uint32_t> term_col;
    std::vector<uint64_t> term_coef;
    std::vector<ColRec> cols;
    std::vector<uint32_t> col_ids;     // lists of ColRec ids (entry columns)
    std::vector<uint32_t> prod_ids;    // pairs of ColRec ids
    std::vector<uint32_t> const_ids;   // ColRec ids
    std::vector<FilterRec> filters;
    std::vector<EntryRec> entries;
    std::vector<CtlZRec> ctl_zs;
    std::vector<LookupRec> lookups;

    uint32_t add_column(const Column& c) {
        ColRec r;
        r.lin_begin = (uint32_t)term_col.size();
        for (auto& p : c.lin) { term_col.push_back(p.first); term_coef.push_back(canon(p.second)); }
        r.lin_end = r.next_begin = (uint32_t)term_col.size();
        for (auto& p : c.next) { term_col.push_back(p.first); term_coef.push_back(canon(p.second)); }
        r.next_end = (uint32_t)term_col.size();
        r.constant = canon(c.constant);
        cols.push_back(r);
        return (uint32_t)cols.size() - 1;
    }
    uint32_t add_filter(const Filter& f) {
        FilterRec r;
        r.prod_begin = (uint32_t)prod_ids.size();
        for (auto& pr : f.products) { uint32_t a = add_column(pr.first), b = add_column(pr.second); prod_ids.push_back(a); prod_ids.push_back(b); }
        r.prod_end = (uint32_t)prod_ids.size();
        r.const_begin = (uint32_t)const_ids.size();
        for (auto& c : f.constants) const_ids.push_back(add_column(c));
        r.const_end = (uint32_t)const_ids.size();
        filters.push_back(r);
        return (uint32_t)filters.size() - 1;
    }
    uint32_t add_entry(const std::vector<Column>& columns, const Filter& f) {
        EntryRec e;
        std::vector<uint32_t> ids;
        for (auto& c : columns) ids.push_back(add_column(c));
        e.col_begin = (uint32_t)col_ids.size();
        for (uint32_t id : ids) col_ids.push_back(id);
        e.col_end = (uint32_t)col_ids.size();
        e.filter = add_filter(f);
        entries.push_back(e);
        return (uint32_t)entries.size() - 1;
    }
};

}  // namespace zkstark
