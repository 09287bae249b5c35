/* tests/adapter_mock/c/piqp.h -- MOCK of the reference's C interface types for the GPU box, where /root/reference does not exist.
 * Only what integration/piqp_batched.c touches: the data / result / status types of interfaces/c/include/piqp_typedef.h:27-175.
 * piqp_settings / piqp_info are taken from include/piqp_b200.h (layout-identical by construction; in the build container
 * tests/test_adapter_header.py compiles piqp_batched.c against the REAL reference headers, whose static asserts prove it). */
#ifndef PIQP_H
#define PIQP_H
#include "piqp_b200.h"
typedef double piqp_float;
typedef int piqp_int;
typedef struct { piqp_int m, n, nnz; piqp_int* p; piqp_int* i; piqp_float* x; } piqp_csc;
typedef struct { piqp_int n, p, m; piqp_float *P, *c, *A, *b, *G, *h_l, *h_u, *x_l, *x_u; } piqp_data_dense;
typedef struct { piqp_int n, p, m; piqp_csc* P; piqp_float* c; piqp_csc* A; piqp_float* b; piqp_csc* G; piqp_float *h_l, *h_u, *x_l, *x_u; } piqp_data_sparse;
typedef b200qp_settings piqp_settings;
typedef enum { PIQP_SOLVED = 1, PIQP_MAX_ITER_REACHED = -1, PIQP_PRIMAL_INFEASIBLE = -2, PIQP_DUAL_INFEASIBLE = -3, PIQP_NUMERICS = -8, PIQP_UNSOLVED = -9,
               PIQP_INVALID_SETTINGS = -10 } piqp_status;
typedef b200qp_info piqp_info;
typedef struct { const piqp_float *x, *y, *z_l, *z_u, *z_bl, *z_bu, *s_l, *s_u, *s_bl, *s_bu; piqp_info info; } piqp_result;
typedef struct { piqp_int is_dense, n, p, m; } piqp_solver_info;
#define PIQP_INF 1e30
#endif
