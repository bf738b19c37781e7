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
// Per-table view of the cross-table lookups: the list of CtlZData a table carries, in the order starky's
// cross_table_lookup_data pushes them (for each CTL, for each challenge: one item per run of consecutive looking
// entries on this table, then the looked item if this table is the looked one).
// ---------------------------------------------------------------------------------------------------------------
struct CtlZItem {
    std::vector<std::pair<std::vector<Column>, Filter>> entries;
    uint32_t challenge = 0;
    uint32_t ctl_index = 0;
    bool looked = false;
    // partial_sums keeps the helper columns only when there is more than one (columns, filter) pair
    uint32_t num_helpers(unsigned constraint_degree) const {
        uint32_t d = constraint_degree - 1;
        return entries.size() > 1 ? (uint32_t)((entries.size() + d - 1) / d) : 0;
    }
};
inline std::vector<CtlZItem> table_ctl_items(uint32_t table, const std::vector<CrossTableLookup>& ctls, unsigned num_challenges) {
    std::vector<CtlZItem> items;
    for (size_t ci = 0; ci < ctls.size(); ci++) {
        const CrossTableLookup& ctl = ctls[ci];
        for (unsigned ch = 0; ch < num_challenges; ch++) {
            size_t i = 0;
            while (i < ctl.looking_tables.size()) {
                size_t j = i;
                while (j < ctl.looking_tables.size() && ctl.looking_tables[j].table == ctl.looking_tables[i].table) j++;
                if (ctl.looking_tables[i].table == table) {
                    CtlZItem it; it.challenge = ch; it.ctl_index = (uint32_t)ci;
                    for (size_t k = i; k < j; k++) it.entries.push_back({ctl.looking_tables[k].columns, ctl.looking_tables[k].filter});
                    items.push_back(it);
                }
                i = j;
            }
            if (ctl.looked_table.table == table) {
                CtlZItem it; it.challenge = ch; it.ctl_index = (uint32_t)ci; it.looked = true;
                it.entries.push_back({ctl.looked_table.columns, ctl.looked_table.filter});
                items.push_back(it);
            }
        }
    }
    return items;
}

// ---------------------------------------------------------------------------------------------------------------
// Flat (POD) form: what the CUDA kernels interpret.  All index ranges are half-open.
// ---------------------------------------------------------------------------------------------------------------
static const uint32_t FLAT_CELL = 0x80000000u;
struct ColRec { uint32_t lin_begin, lin_end, next_begin, next_end; uint64_t constant; };
struct FilterRec { uint32_t prod_begin, prod_end;     // prod_ids[2k], prod_ids[2k+1] are ColRec ids
                   uint32_t const_begin, const_end; };   // const_ids[k] are ColRec ids
// one (columns, filter) pair: ColRec ids col_ids[col_begin .. col_end)
struct EntryRec { uint32_t col_begin, col_end; uint32_t filter; uint32_t pad; };
// one CtlZData of a table: entries [entry_begin, entry_end); helper columns live at aux index helper_begin..+num_helpers,
// the running sum Z at aux index z_col
struct CtlZRec { uint32_t entry_begin, entry_end; uint32_t num_helpers; uint32_t challenge; uint32_t helper_begin; uint32_t z_col;
                 // position of this item's first constraint inside the table's CTL section (num_helpers + 2 constraints, or 2), and
                 // the index of the item that differs from this one only by the challenge (NO_TWIN if there is none)
                 uint32_t cons_begin; uint32_t twin; };
static const uint32_t NO_TWIN = 0xFFFFFFFFu;
// one in-table Lookup for one challenge: single-column entries [entry_begin, entry_end); helpers at aux index
// helper_begin..+num_helpers, Z at z_col
struct LookupRec { uint32_t entry_begin, entry_end; uint32_t table_col, freq_col; uint32_t num_helpers; uint32_t challenge;
                   uint32_t helper_begin; uint32_t z_col; };

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
    uint32_t num_lookup_cols = 0, num_ctl_helpers = 0, num_ctl_zs = 0;
    uint32_t ctl_num_constraints = 0;   // constraints of the CTL section
    uint32_t ctl_paired = 0;            // 1: two challenges and every item has its twin -> the evaluators share the column walks
    uint32_t num_aux() const { return num_lookup_cols + num_ctl_helpers + num_ctl_zs; }

    // Column id: most CTL / lookup columns are a plain cell of the local row (coefficient 1, no constant); those are encoded in
    // the id itself (FLAT_CELL | column index) and cost the evaluators one load instead of a descriptor walk
    uint32_t add_column(const Column& c) {
        if (c.lin.size() == 1 && c.next.empty() && canon(c.constant) == 0 && canon(c.lin[0].second) == 1 && c.lin[0].first < FLAT_CELL)
            return FLAT_CELL | c.lin[0].first;
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
        std::vector<uint32_t> pp, cc;
        for (auto& pr : f.products) { pp.push_back(add_column(pr.first)); pp.push_back(add_column(pr.second)); }
        for (auto& c : f.constants) cc.push_back(add_column(c));
        r.prod_begin = (uint32_t)prod_ids.size();
        prod_ids.insert(prod_ids.end(), pp.begin(), pp.end());
        r.prod_end = (uint32_t)prod_ids.size();
        r.const_begin = (uint32_t)const_ids.size();
        const_ids.insert(const_ids.end(), cc.begin(), cc.end());
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
        e.pad = 0;
        entries.push_back(e);
        return (uint32_t)entries.size() - 1;
    }
};

// Auxiliary-column layout of one table (starky prove_with_commitment): lookup columns (per lookup, per challenge:
// helpers..., Z) ++ CTL helper columns of every item ++ CTL Z of every item.
inline Flat build_table_flat(const std::vector<Lookup>& lookups, const std::vector<CtlZItem>& items, unsigned num_challenges,
                             unsigned constraint_degree) {
    Flat f;
    uint32_t aux = 0;
    for (const Lookup& l : lookups) {
        uint32_t first_entry = (uint32_t)f.entries.size();
        for (size_t i = 0; i < l.columns.size(); i++) f.add_entry({l.columns[i]}, l.filter_columns[i]);
        uint32_t last_entry = (uint32_t)f.entries.size();
        uint32_t tcol = f.add_column(l.table_column), fcol = f.add_column(l.frequencies_column);
        uint32_t nh = (uint32_t)l.num_helper_columns(constraint_degree) - 1;
        for (unsigned ch = 0; ch < num_challenges; ch++) {
            LookupRec r;
            r.entry_begin = first_entry; r.entry_end = last_entry; r.table_col = tcol; r.freq_col = fcol;
            r.num_helpers = nh; r.challenge = ch; r.helper_begin = aux; r.z_col = aux + nh;
            aux += nh + 1;
            f.lookups.push_back(r);
        }
    }
    f.num_lookup_cols = aux;
    uint32_t helpers = 0;
    for (const CtlZItem& it : items) helpers += it.num_helpers(constraint_degree);
    f.num_ctl_helpers = helpers;
    f.num_ctl_zs = (uint32_t)items.size();
    uint32_t hpos = aux, zpos = aux + helpers;
    for (const CtlZItem& it : items) {
        CtlZRec r;
        r.entry_begin = (uint32_t)f.entries.size();
        for (auto& e : it.entries) f.add_entry(e.first, e.second);
        r.entry_end = (uint32_t)f.entries.size();
        r.num_helpers = it.num_helpers(constraint_degree);
        r.challenge = it.challenge;
        r.helper_begin = hpos; hpos += r.num_helpers;
        r.z_col = zpos++;
        r.cons_begin = f.ctl_num_constraints;
        f.ctl_num_constraints += r.num_helpers ? r.num_helpers + 2 : 2;
        r.twin = NO_TWIN;
        f.ctl_zs.push_back(r);
    }
    // twins: same CTL, same position inside the (CTL, challenge) group, other challenge
    if (num_challenges == 2) {
        f.ctl_paired = 1;
        for (size_t i = 0; i < items.size(); i++) {
            if (items[i].challenge != 0) continue;
            size_t ord = 0;
            for (size_t k = 0; k < i; k++) if (items[k].ctl_index == items[i].ctl_index && items[k].challenge == 0) ord++;
            size_t seen = 0;
            for (size_t k = 0; k < items.size(); k++) {
                if (items[k].ctl_index != items[i].ctl_index || items[k].challenge != 1) continue;
                if (seen++ == ord) {
                    if (items[k].entries.size() == items[i].entries.size() && items[k].looked == items[i].looked) {
                        f.ctl_zs[i].twin = (uint32_t)k; f.ctl_zs[k].twin = (uint32_t)i;
                    }
                    break;
                }
            }
            if (f.ctl_zs[i].twin == NO_TWIN) f.ctl_paired = 0;
        }
        for (const CtlZRec& r : f.ctl_zs) if (r.twin == NO_TWIN) f.ctl_paired = 0;
    }
    return f;
}

// raw-pointer view of a Flat (host vectors or device copies)
struct FlatView {
    const uint32_t* term_col; const uint64_t* term_coef; const ColRec* cols; const uint32_t* col_ids;
    const uint32_t* prod_ids; const uint32_t* const_ids; const FilterRec* filters; const EntryRec* entries;
    const CtlZRec* ctl_zs; const LookupRec* lookups;
    uint32_t n_ctl_zs, n_lookups, num_lookup_cols, num_ctl_helpers, num_ctl_zs;
    uint32_t ctl_num_constraints, ctl_paired;
};

}  // namespace zkstark
