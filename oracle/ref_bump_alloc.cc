/*
 * oracle/ref_bump_alloc.cc -- deterministic heap addresses for the reference's quad-tree tie-break.
 *
 * TEST INFRASTRUCTURE ONLY.  ORBextractor::DistributeOctTree sorts pair<int, ExtractorNode*> (ORBextractor.cc:684), so
 * nodes holding the same number of keys are ordered by HEAP ADDRESS: the reference's output depends on the allocator.
 * To compare the reference's object code with the restatement, this file replaces operator new/delete INSIDE
 * libref_orbextractor.so (linked -Bsymbolic): while ref_bump_enable(1) is active, allocations of exactly
 * sizeof(std::list<ExtractorNode> node) come from an arena that hands out ascending addresses and never reuses one, i.e.
 * "created later <=> higher address" -- the tie-break orb_oracle.c and the CUDA kernel define.  With
 * ref_bump_enable(0) everything goes to malloc (glibc address reuse, whatever order that gives).
 */
#include <cstdio>
#include <cstdlib>
#include <list>
#include <new>
#include "ORBextractor.h"

namespace {
struct NodeProbe { void *a, *b; iORB_SLAM::ExtractorNode n; };          /* layout of std::_List_node<ExtractorNode> */
const size_t kNodeBytes = sizeof(NodeProbe);
const size_t kArenaBytes = (size_t)256 << 20;
char *g_arena = nullptr;
size_t g_off = 0;
long g_live = 0;
int g_on = 0;
}

extern "C" void ref_bump_enable(int on) { g_on = on; }
extern "C" void ref_bump_reset(void) { if (g_live == 0) g_off = 0; }   /* called at the start of every extraction */
extern "C" long ref_bump_nodes(void) { return (long)(g_off / ((kNodeBytes + 15) & ~(size_t)15)); }

void *operator new(size_t n)
{
    if (g_on && n == kNodeBytes) {
        if (!g_arena) g_arena = (char *)std::malloc(kArenaBytes);
        const size_t sz = (n + 15) & ~(size_t)15;
        if (g_arena && g_off + sz <= kArenaBytes) { void *p = g_arena + g_off; g_off += sz; g_live++; return p; }
    }
    void *p = std::malloc(n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
void operator delete(void *p) noexcept
{
    if (g_arena && (char *)p >= g_arena && (char *)p < g_arena + kArenaBytes) { g_live--; return; }
    std::free(p);
}
void operator delete(void *p, size_t) noexcept { operator delete(p); }
void *operator new[](size_t n) { return operator new(n); }
void operator delete[](void *p) noexcept { operator delete(p); }
void operator delete[](void *p, size_t) noexcept { operator delete(p); }
