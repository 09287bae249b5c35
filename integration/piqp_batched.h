/*
 * integration/piqp_batched.h -- REFERENCE-SIDE binding of the batched path: the reference's public C API
 * (interfaces/c/include/piqp.h:21-43) with a leading batch dimension, implemented over libpiqp_b200's b200qp_* entry points.
 *
 * A maintainer adds this header + piqp_batched.c to interfaces/c/ and links -lpiqp_b200; nothing in piqp.h changes.  It uses the
 * reference's OWN types: piqp_data_dense / piqp_data_sparse (arrays carry a leading batch axis, instance-major), piqp_settings,
 * piqp_info, piqp_result, piqp_status (interfaces/c/include/piqp_typedef.h:27-175).  b200qp_settings / b200qp_info are
 * layout-identical to piqp_settings / piqp_info by construction; piqp_batched.c checks that at compile time.
 *
 *   piqp_batched_workspace* w;
 *   piqp_setup_dense_batched(&w, batch, &data, &settings);     // data.P = [batch][n][n] row-major, data.c = [batch][n], ...
 *   piqp_solve_batched(w);                                     // all instances, device-resident IP loop
 *   w->result[b].x, w->result[b].info.status, ...              // per-instance piqp_result, as after piqp_solve
 *   piqp_update_dense_batched(w, P, c, A, b, G, h_l, h_u, x_l, x_u);   // NULL = keep, like piqp_update_dense
 *   piqp_cleanup_batched(w);
 */
#ifndef PIQP_BATCHED_H
#define PIQP_BATCHED_H

#include "piqp.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    void* handle;                 /* b200qp_handle                                                        */
    piqp_int batch;
    piqp_solver_info solver_info; /* is_dense, n, p, m -- as in piqp_workspace                             */
    piqp_result* result;          /* [batch]; pointers into the contiguous host buffers below              */
    piqp_float* buffers;          /* x|y|z_l|z_u|z_bl|z_bu|s_l|s_u|s_bl|s_bu, each [batch][len]            */
} piqp_batched_workspace;

/* piqp_setup_dense (piqp.h:27) for `batch` QPs of one shape; every array of `data` has a leading batch axis */
piqp_int piqp_setup_dense_batched(piqp_batched_workspace** workspace, piqp_int batch, const piqp_data_dense* data, const piqp_settings* settings);
/* piqp_setup_sparse (piqp.h:28): the instances share the CSC patterns of data->P / A / G; their x arrays are [batch][nnz] */
piqp_int piqp_setup_sparse_batched(piqp_batched_workspace** workspace, piqp_int batch, const piqp_data_sparse* data, const piqp_settings* settings);
/* piqp_update_settings / piqp_update_dense / piqp_update_sparse (piqp.h:30-40) */
piqp_int piqp_update_settings_batched(piqp_batched_workspace* workspace, const piqp_settings* settings);
piqp_int piqp_update_dense_batched(piqp_batched_workspace* workspace, piqp_float* P, piqp_float* c, piqp_float* A, piqp_float* b,
                                   piqp_float* G, piqp_float* h_l, piqp_float* h_u, piqp_float* x_l, piqp_float* x_u);
piqp_int piqp_update_sparse_batched(piqp_batched_workspace* workspace, piqp_float* Px, piqp_float* c, piqp_float* Ax, piqp_float* b,
                                    piqp_float* Gx, piqp_float* h_l, piqp_float* h_u, piqp_float* x_l, piqp_float* x_u);
/* piqp_solve (piqp.h:42): returns PIQP_SOLVED if every instance is solved, else the status of the first instance that is not */
piqp_status piqp_solve_batched(piqp_batched_workspace* workspace);
void piqp_cleanup_batched(piqp_batched_workspace* workspace);

#ifdef __cplusplus
}
#endif
#endif /* PIQP_BATCHED_H */
