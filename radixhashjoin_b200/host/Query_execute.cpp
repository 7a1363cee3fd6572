// Query_execute.cpp -- drop-in for ONE member function of the reference's Query.cpp:
//     void Query::execute(JobScheduler &js, std::vector<relList> &relations)          (Query.h:50, Query.cpp:204-211)
// called unchanged from QueryJob::run (MainScheduler.cpp:23-26) on up to NUM_OF_THREADS query threads.
//
// The reference's execute = run_filters (unordered_sets of row ids) -> run_joins (create_relation, the join, Result
// pages, update_intermediate) -> column_proj, everything on the host.  This one hands the parsed query to
// rhj_query_execute (include/rhj.h): row-id lists, relations, join results and the intermediate stay in HBM from the
// first filter to the last checksum; the relList columns are uploaded once per process; per query only a few counts and
// the checksums cross PCIe.  Parsing (Query::Query, read_*), printing (Query::print) and the schedulers are the
// reference's own, unmodified: the build compiles the reference's Query.cpp with -Dexecute=execute_on_cpu so that its
// definition of this one function does not collide (see the Makefile; a maintainer would simply replace the body).
// No CPU fallback: a failing library call ends the process with the library's message.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "Query.h"   // the reference's header
#include "rhj.h"
#include "thread_ctx.h"

void Query::execute(JobScheduler &, std::vector<relList> &relations) {
    rhj_host::Scope timer(0);
    std::vector<rhj_q_relation> binds(table.size());
    for (size_t b = 0; b < table.size(); b++) {
        const relList &rel = relations[table[b]];
        binds[b].columns = rel.values;
        binds[b].num_tuples = rel.num_tuples;
        binds[b].num_columns = rel.num_columns;
    }
    std::vector<rhj_q_filter> fl(filter.size());
    for (size_t i = 0; i < filter.size(); i++) {
        fl[i].binding = (uint32_t) filter[i].table;
        fl[i].column = (uint32_t) filter[i].column;
        fl[i].op = filter[i].op;
        fl[i].reserved0 = 0;
        fl[i].constant = filter[i].number;
    }
    std::vector<rhj_q_join> jn(join.size());
    for (size_t i = 0; i < join.size(); i++) {
        jn[i].binding1 = (uint32_t) join[i].table1;
        jn[i].column1 = (uint32_t) join[i].column1;
        jn[i].binding2 = (uint32_t) join[i].table2;
        jn[i].column2 = (uint32_t) join[i].column2;
    }
    std::vector<rhj_q_proj> pj(proj.size());
    for (size_t i = 0; i < proj.size(); i++) {
        pj[i].binding = (uint32_t) proj[i].table;
        pj[i].column = (uint32_t) proj[i].column;
    }
    rhj_query_desc d;
    d.n_bindings = (uint32_t) binds.size();
    d.n_filters = (uint32_t) fl.size();
    d.n_joins = (uint32_t) jn.size();
    d.n_projs = (uint32_t) pj.size();
    static const bool keep_order = getenv("RHJ_KEEP_JOIN_ORDER") != nullptr;   // default: cheapest-first (same checksums)
    d.reorder_joins = keep_order ? 0 : 1;
    d.reserved0 = 0;
    d.bindings = binds.data();
    d.filters = fl.data();
    d.joins = jn.data();
    d.projs = pj.data();
    std::vector<uint64_t> sums(proj.size() + 1, 0);
    int empty = 0;
    rhj_query_stats st;
    rhj_host::check(rhj_query_execute(rhj_host::thread_ctx(), &d, sums.data(), &empty, &st), "rhj_query_execute");
    rhj_host::pcie().add(st);
    filtered_out = empty != 0;
    for (size_t i = 0; i < proj.size(); i++) proj[i].sum = sums[i];
}
