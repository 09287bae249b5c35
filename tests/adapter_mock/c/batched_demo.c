/* the reference's own known-answer QP (tests/src/c_interface_test.cpp / dense/solver_test.cpp:30-67) through the reference-side
 * batched binding integration/piqp_batched.c: two instances, dense and sparse entry points */
#include <math.h>
#include <stdio.h>
#include "piqp_batched.h"
int main(void) {
    double P[2][4] = {{6, 0, 0, 4}, {6, 0, 0, 4}}, c[2][2] = {{-1, -4}, {-1, -4}}, A[2][2] = {{1, -2}, {1, -2}}, b[2][1] = {{0}, {0}};
    double G[2][6] = {{1, 0, 1, 0, 1, 0}, {1, 0, 1, 0, 1, 0}}, h_l[2][3] = {{-1, -PIQP_INF, -2}, {-1, -PIQP_INF, -2}}, h_u[2][3] = {{PIQP_INF, 1, 2}, {PIQP_INF, 1, 2}};
    double x_l[2][2] = {{-PIQP_INF, -1}, {-PIQP_INF, -1}}, x_u[2][2] = {{PIQP_INF, 1}, {PIQP_INF, 1}};
    piqp_data_dense d = {2, 1, 3, &P[0][0], &c[0][0], &A[0][0], &b[0][0], &G[0][0], &h_l[0][0], &h_u[0][0], &x_l[0][0], &x_u[0][0]};
    piqp_settings st;
    piqp_batched_workspace* w = NULL;
    int k, bad = 0;
    b200qp_set_default_settings_dense(&st);
    if (piqp_setup_dense_batched(&w, 2, &d, &st) != 0) { printf("setup failed: %s\n", b200_last_error()); return 2; }
    if (piqp_solve_batched(w) != PIQP_SOLVED) bad = 1;
    for (k = 0; k < 2; k++) {
        printf("dense  instance %d: status %d iter %d x = (%.7f, %.7f) y = %.7f\n", k, (int)w->result[k].info.status, w->result[k].info.iter, w->result[k].x[0], w->result[k].x[1], w->result[k].y[0]);
        if (fabs(w->result[k].x[0] - 0.4285714) > 1e-6 || fabs(w->result[k].x[1] - 0.2142857) > 1e-6 || fabs(w->result[k].y[0] + 1.5714286) > 1e-6) bad = 1;
    }
    piqp_cleanup_batched(w);
    {   /* sparse twin: CSC of P (upper), A, G; value arrays [batch][nnz] */
        int Pp[3] = {0, 1, 2}, Pi[2] = {0, 1}, Ap[3] = {0, 1, 2}, Ai[2] = {0, 0}, Gp[3] = {0, 3, 3}, Gi[3] = {0, 1, 2};
        double Px[2][2] = {{6, 4}, {6, 4}}, Ax[2][2] = {{1, -2}, {1, -2}}, Gx[2][3] = {{1, 1, 1}, {1, 1, 1}};
        piqp_csc Pm = {2, 2, 2, Pp, Pi, &Px[0][0]}, Am = {1, 2, 2, Ap, Ai, &Ax[0][0]}, Gm = {3, 2, 3, Gp, Gi, &Gx[0][0]};
        piqp_data_sparse s = {2, 1, 3, &Pm, &c[0][0], &Am, &b[0][0], &Gm, &h_l[0][0], &h_u[0][0], &x_l[0][0], &x_u[0][0]};
        b200qp_set_default_settings_sparse(&st);
        if (piqp_setup_sparse_batched(&w, 2, &s, &st) != 0) { printf("sparse setup failed: %s\n", b200_last_error()); return 2; }
        if (piqp_solve_batched(w) != PIQP_SOLVED) bad = 1;
        for (k = 0; k < 2; k++) {
            printf("sparse instance %d: status %d iter %d x = (%.7f, %.7f)\n", k, (int)w->result[k].info.status, w->result[k].info.iter, w->result[k].x[0], w->result[k].x[1]);
            if (fabs(w->result[k].x[0] - 0.4285714) > 1e-6 || fabs(w->result[k].x[1] - 0.2142857) > 1e-6) bad = 1;
        }
        piqp_cleanup_batched(w);
    }
    printf("%s\n", bad ? "BATCHED_BINDING_FAIL" : "BATCHED_BINDING_OK");
    return bad;
}
