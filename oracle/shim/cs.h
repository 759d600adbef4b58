// Test-infrastructure shim (oracle/Makefile, oracle/_ref only): lets UNMODIFIED reference sources compile without SuiteSparse / CXSparse (graphSlam6D.cc:317-372,399-419 call six cs_* functions; dense implementations in oracle/shim_impl.cc).
#pragma once
struct cs_sparse;
typedef struct cs_sparse cs;
extern "C" {
cs* cs_spalloc(int m, int n, int nzmax, int values, int triplet);
int cs_entry(cs* T, int i, int j, double x);
cs* cs_compress(const cs* T);
int cs_dropzeros(cs* A);
int cs_cholsol(int order, const cs* A, double* b);
int cs_qrsol(int order, const cs* A, double* b);
cs* cs_spfree(cs* A);
int cs_print(const cs* A, int brief);
}
