/* integration/piqp_batched.c -- see piqp_batched.h.  C99; needs the reference's interfaces/c/include and include/piqp_b200.h. */
#include "piqp_batched.h"

#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#include "piqp_b200.h"

/* the two structs travel through the ABI by pointer cast: prove the layouts agree */
#define PIQP_B_STATIC_ASSERT(c, name) typedef char piqp_b_static_assert_##name[(c) ? 1 : -1]
PIQP_B_STATIC_ASSERT(sizeof(piqp_settings) == sizeof(b200qp_settings), settings_size);
PIQP_B_STATIC_ASSERT(offsetof(piqp_settings, kkt_solver) == offsetof(b200qp_settings, kkt_solver), settings_kkt_solver);
PIQP_B_STATIC_ASSERT(offsetof(piqp_settings, compute_timings) == offsetof(b200qp_settings, compute_timings), settings_tail);
PIQP_B_STATIC_ASSERT(sizeof(piqp_info) == sizeof(b200qp_info), info_size);
PIQP_B_STATIC_ASSERT(offsetof(piqp_info, primal_obj) == offsetof(b200qp_info, primal_obj), info_obj);
PIQP_B_STATIC_ASSERT(offsetof(piqp_info, run_time) == offsetof(b200qp_info, run_time), info_tail);
PIQP_B_STATIC_ASSERT(sizeof(piqp_float) == sizeof(double) && sizeof(piqp_int) == sizeof(int), fp64_int32);

static piqp_batched_workspace* make_workspace(void* handle, piqp_int batch, piqp_int is_dense, piqp_int n, piqp_int p, piqp_int m) {
    piqp_batched_workspace* w = (piqp_batched_workspace*)calloc(1, sizeof *w);
    size_t per = (size_t)(5 * n + p + 4 * m), off = 0, B = (size_t)batch;
    piqp_int k;
    if (!w) return NULL;
    w->handle = handle; w->batch = batch;
    w->solver_info.is_dense = is_dense; w->solver_info.n = n; w->solver_info.p = p; w->solver_info.m = m;
    w->result = (piqp_result*)calloc(B, sizeof(piqp_result));
    w->buffers = (piqp_float*)calloc(B * per + 1, sizeof(piqp_float));
    if (!w->result || !w->buffers) { free(w->result); free(w->buffers); free(w); return NULL; }
    {   /* x|y|z_l|z_u|z_bl|z_bu|s_l|s_u|s_bl|s_bu, each [batch][len] */
        const size_t len[10] = {(size_t)n, (size_t)p, (size_t)m, (size_t)m, (size_t)n, (size_t)n, (size_t)m, (size_t)m, (size_t)n, (size_t)n};
        const piqp_float* base[10];
        int f;
        for (f = 0; f < 10; f++) { base[f] = w->buffers + off; off += B * len[f]; }
        for (k = 0; k < batch; k++) {
            piqp_result* r = &w->result[k];
            r->x = base[0] + (size_t)k * len[0]; r->y = base[1] + (size_t)k * len[1]; r->z_l = base[2] + (size_t)k * len[2]; r->z_u = base[3] + (size_t)k * len[3];
            r->z_bl = base[4] + (size_t)k * len[4]; r->z_bu = base[5] + (size_t)k * len[5]; r->s_l = base[6] + (size_t)k * len[6]; r->s_u = base[7] + (size_t)k * len[7];
            r->s_bl = base[8] + (size_t)k * len[8]; r->s_bu = base[9] + (size_t)k * len[9];
            r->info.status = PIQP_UNSOLVED;
        }
    }
    return w;
}

piqp_int piqp_setup_dense_batched(piqp_batched_workspace** workspace, piqp_int batch, const piqp_data_dense* d, const piqp_settings* settings) {
    b200qp_handle* h = NULL;
    int rc = b200qp_setup_dense(&h, batch, d->n, d->p, d->m, d->P, d->c, d->A, d->b, d->G, d->h_l, d->h_u, d->x_l, d->x_u,
                                (const b200qp_settings*)settings, /*device=*/0, /*on_device=*/0);
    if (rc != B200_OK) return rc;
    *workspace = make_workspace(h, batch, 1, d->n, d->p, d->m);
    return *workspace ? 0 : -1;
}

piqp_int piqp_setup_sparse_batched(piqp_batched_workspace** workspace, piqp_int batch, const piqp_data_sparse* d, const piqp_settings* settings) {
    b200qp_handle* h = NULL;
    int rc = b200qp_setup_sparse(&h, batch, d->n, d->p, d->m, d->P->p, d->P->i, d->P->x, d->c,
                                 d->A ? d->A->p : NULL, d->A ? d->A->i : NULL, d->A ? d->A->x : NULL, d->b,
                                 d->G ? d->G->p : NULL, d->G ? d->G->i : NULL, d->G ? d->G->x : NULL, d->h_l, d->h_u, d->x_l, d->x_u,
                                 (const b200qp_settings*)settings, 0, 0);
    if (rc != B200_OK) return rc;
    *workspace = make_workspace(h, batch, 0, d->n, d->p, d->m);
    return *workspace ? 0 : -1;
}

piqp_int piqp_update_settings_batched(piqp_batched_workspace* w, const piqp_settings* settings) {
    return b200qp_update_settings((b200qp_handle*)w->handle, (const b200qp_settings*)settings);
}
piqp_int piqp_update_dense_batched(piqp_batched_workspace* w, piqp_float* P, piqp_float* c, piqp_float* A, piqp_float* b, piqp_float* G,
                                   piqp_float* h_l, piqp_float* h_u, piqp_float* x_l, piqp_float* x_u) {
    return b200qp_update_dense((b200qp_handle*)w->handle, P, c, A, b, G, h_l, h_u, x_l, x_u, 0);
}
piqp_int piqp_update_sparse_batched(piqp_batched_workspace* w, piqp_float* Px, piqp_float* c, piqp_float* Ax, piqp_float* b, piqp_float* Gx,
                                    piqp_float* h_l, piqp_float* h_u, piqp_float* x_l, piqp_float* x_u) {
    return b200qp_update_sparse((b200qp_handle*)w->handle, Px, c, Ax, b, Gx, h_l, h_u, x_l, x_u, 0);
}

piqp_status piqp_solve_batched(piqp_batched_workspace* w) {
    b200qp_handle* h = (b200qp_handle*)w->handle;
    const size_t B = (size_t)w->batch, n = (size_t)w->solver_info.n, p = (size_t)w->solver_info.p, m = (size_t)w->solver_info.m;
    piqp_float* x = w->buffers;
    piqp_float *y = x + B * n, *z_l = y + B * p, *z_u = z_l + B * m, *z_bl = z_u + B * m, *z_bu = z_bl + B * n, *s_l = z_bu + B * n, *s_u = s_l + B * m,
               *s_bl = s_u + B * m, *s_bu = s_bl + B * n;
    b200qp_info* infos;
    piqp_status worst = PIQP_SOLVED;
    piqp_int k;
    if (b200qp_solve(h) != B200_OK) return PIQP_NUMERICS;
    if (b200qp_get_result(h, x, y, z_l, z_u, z_bl, z_bu, s_l, s_u, s_bl, s_bu, 0) != B200_OK) return PIQP_NUMERICS;
    infos = (b200qp_info*)malloc(B * sizeof *infos);
    if (!infos || b200qp_get_info(h, infos) != B200_OK) { free(infos); return PIQP_NUMERICS; }
    for (k = 0; k < w->batch; k++) {
        memcpy(&w->result[k].info, &infos[k], sizeof(piqp_info));
        if (worst == PIQP_SOLVED && w->result[k].info.status != PIQP_SOLVED) worst = w->result[k].info.status;
    }
    free(infos);
    return worst;
}

void piqp_cleanup_batched(piqp_batched_workspace* w) {
    if (!w) return;
    b200qp_cleanup((b200qp_handle*)w->handle);
    free(w->result); free(w->buffers); free(w);
}
