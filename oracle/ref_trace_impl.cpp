// ref_trace_impl.cpp -- TEST INFRASTRUCTURE ONLY (see oracle/Makefile, target join_ref_trace).
//
// Compiles the reference's own Result.cpp (from /root/reference, via -I) with
// its join entry point renamed, so that ref_trace_hook.cpp can define
// Result::multiRadixHashJoin as "call the reference, then log a digest of the
// inputs and of the result".  This is how tests/golden/small_joins.json was made.
#define multiRadixHashJoin multiRadixHashJoin_reference
#include "Result.cpp"
#undef multiRadixHashJoin

extern "C" void rhj_trace_call_reference(void *res, void *js, void *relR, void *relS) {
    ((Result *) res)->multiRadixHashJoin_reference(*(JobScheduler *) js, *(relation *) relR, *(relation *) relS);
}
