// Test-infrastructure shim (oracle/Makefile, oracle/_ref only): lets UNMODIFIED reference sources compile without Boost.Graph (graph.h:18-26 only names the adjacency_list type; ELCH is out of scope).
#pragma once
// shim: graph.h only names the type (ELCH's loop graph lives in elch6D.cc, out of scope)
namespace boost {
struct listS {}; struct vecS {}; struct undirectedS {}; struct no_property {};
enum edge_weight_t { edge_weight };
template <class Tag, class T> struct property {};
template <class A, class B, class C, class D, class E> class adjacency_list {};
}
