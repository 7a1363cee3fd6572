// intermediate.cpp -- drop-in replacement for the reference's intermediate.cpp.
//
// Same two entry points, same signatures (intermediate.h:10-14), called unchanged from
// Query::run_joins (Query.cpp:168-170,193).  What changes is the algorithm of update_intermediate:
// the reference re-scans the whole intermediate once per result pair (change_intermediate /
// change_both_intermediate, intermediate.cpp:52-87: O(pairs x rows), 99 % of the small.work wall
// time), here
//   case 1 (neither binding joined before, 153-161): the pairs are unzipped into two columns;
//   case 2 (one binding joined before, 162-170):  an N:M equi-join between the result pairs and
//          the existing row-id column + gathers            -> rhj_intermediate_expand_host;
//   case 3 (both joined before, 171-180):         a semi-join on the (row id, row id) pair
//                                                           -> rhj_intermediate_filter_host;
// both on the GPU through the C ABI.  The new intermediate holds the same multiset of rows as the
// reference's; row order differs, which no consumer observes (create_relation de-duplicates row
// ids, structs.cpp:238-241; column_proj sums, Query.cpp:66-74).
#include <cstdio>
#include <cstdlib>

#include "intermediate.h"   // the reference's header
#include "rhj.h"
#include "thread_ctx.h"

using std::unordered_map;
using std::unordered_set;
using std::vector;

namespace {

// The join result as one flat array, in page-walk order (head page holds `size` pairs, every other
// page `capacity`; intermediate.cpp:151-179).
vector<rhj_pair> flatten(const Result &results) {
    size_t n = 0;
    for (bucket_info *pg = results.head; pg != nullptr; pg = pg->next)
        n += (pg == results.head) ? results.size : results.capacity;
    vector<rhj_pair> flat;
    flat.reserve(n);
    for (bucket_info *pg = results.head; pg != nullptr; pg = pg->next) {
        size_t cnt = (pg == results.head) ? results.size : results.capacity;
        auto kt = (const rhj_pair *) &pg[1];
        flat.insert(flat.end(), kt, kt + cnt);
    }
    return flat;
}

}  // namespace

// Same-binding predicate t.a = t.b (Query.cpp:168-170).  First branch as in the reference
// (intermediate.cpp:16-25).  The reference's second branch (26-42) dereferences end() and erases
// with a foreign iterator; it is unreachable in the pinned workload, and is given its evident
// intent here: keep the intermediate rows whose two columns are equal.
void parse_table(join_info &join, relList &relation, unordered_map<uint64_t, unordered_set<uint64_t> > &filtered,
                 vector<vector<uint64_t> > &intermediate) {
    const uint64_t *c1 = relation.values[join.column1];
    const uint64_t *c2 = relation.values[join.column2];
    vector<uint64_t> &mine = intermediate[join.table1];
    if (mine.empty()) {
        auto it = filtered.find(join.table1);
        for (uint64_t rowid : it->second)
            if (c1[rowid] == c2[rowid]) mine.push_back(rowid);
        return;
    }
    vector<size_t> keep;
    for (size_t e = 0; e < mine.size(); e++)
        if (c1[mine[e]] == c2[mine[e]]) keep.push_back(e);
    for (auto &col : intermediate) {
        if (col.empty()) continue;
        vector<uint64_t> next;
        next.reserve(keep.size());
        for (size_t e : keep) next.push_back(col[e]);
        col.swap(next);
    }
}

void update_intermediate(vector<vector<uint64_t> > &intermediate, const Result &results, join_info &join) {
    const size_t nb = intermediate.size();
    vector<rhj_pair> pairs = flatten(results);
    const bool has1 = !intermediate[join.table1].empty();
    const bool has2 = !intermediate[join.table2].empty();

    rhj_host::Scope timer(!has1 && !has2 ? 1 : has1 != has2 ? 2 : 3);
    if (!has1 && !has2) {  // case 1: unzip
        vector<uint64_t> &t1 = intermediate[join.table1];
        vector<uint64_t> &t2 = intermediate[join.table2];
        t1.resize(pairs.size());
        t2.resize(pairs.size());
        for (size_t i = 0; i < pairs.size(); i++) {
            t1[i] = pairs[i].keyR;
            t2[i] = pairs[i].keyS;
        }
        return;
    }

    // live columns of the old intermediate, in binding order
    vector<size_t> live;
    vector<const uint64_t *> cols;
    for (size_t i = 0; i < nb; i++)
        if (!intermediate[i].empty()) {
            live.push_back(i);
            cols.push_back(intermediate[i].data());
        }
    vector<uint64_t *> out(cols.size() + 1, nullptr);
    uint64_t rows = 0;
    rhj_ctx *ctx = rhj_host::thread_ctx();
    vector<vector<uint64_t> > next(nb);

    if (has1 != has2) {  // case 2: expand
        const size_t full = has1 ? join.table1 : join.table2;
        const size_t fresh = has1 ? join.table2 : join.table1;
        // the already-joined binding is the S side of the pair iff it is table2
        const int match_on_S = has1 ? 0 : 1;
        rhj_host::check(rhj_intermediate_expand_host(ctx, intermediate[full].data(), intermediate[full].size(), pairs.data(),
                                                     pairs.size(), match_on_S, cols.data(), (uint32_t) cols.size(),
                                                     out.data(), &rows),
                        "rhj_intermediate_expand_host");
        for (size_t k = 0; k < live.size(); k++) next[live[k]].assign(out[k], out[k] + rows);
        if (rows) next[fresh].assign(out[cols.size()], out[cols.size()] + rows);
    } else {  // case 3: filter
        rhj_host::check(rhj_intermediate_filter_host(ctx, intermediate[join.table1].data(), intermediate[join.table2].data(),
                                                     intermediate[join.table1].size(), pairs.data(), pairs.size(),
                                                     cols.data(), (uint32_t) cols.size(), out.data(), &rows),
                        "rhj_intermediate_filter_host");
        for (size_t k = 0; k < live.size(); k++) next[live[k]].assign(out[k], out[k] + rows);
    }
    intermediate.swap(next);
}
