/* examples/c_api_demo.c -- the C-ABI of libpiqp_b200.so from plain C99 (what a cgo / JNI / ctypes binding sees).
 *
 *   gcc -std=c99 -I include examples/c_api_demo.c -L piqp_b200 -lpiqp_b200 -Wl,-rpath,$PWD/piqp_b200 -o c_api_demo && ./c_api_demo
 *
 * Solves the reference's 2-variable known-answer QP (tests/src/dense/solver_test.cpp:30-73) twice in one batch through the
 * batched interface (b200qp_*, the twin of piqp_setup_dense / piqp_solve).
 * Needs a B200; without a device every call returns a negative code and b200_last_error() says why (there is no CPU path). */
#include <stdio.h>
#include <math.h>
#include "piqp_b200.h"

int main(void) {
    /* min 1/2 x^T P x + c^T x  s.t.  A x = b,  h_l <= G x <= h_u,  x_l <= x <= x_u  (row-major like piqp_data_dense) */
    const int batch = 2, n = 2, p = 1, m = 3;
    const double P[2][4] = {{6, 0, 0, 4}, {6, 0, 0, 4}}, c[2][2] = {{-1, -4}, {-1, -4}};
    const double A[2][2] = {{1, -2}, {1, -2}}, b[2][1] = {{0}, {0}};
    const double G[2][6] = {{1, 0, 1, 0, 1, 0}, {1, 0, 1, 0, 1, 0}};                       /* 3 x 2, row-major */
    const double h_l[2][3] = {{-1, -INFINITY, -2}, {-1, -INFINITY, -2}}, h_u[2][3] = {{INFINITY, 1, 2}, {INFINITY, 1, 2}};
    const double x_l[2][2] = {{-INFINITY, -1}, {-INFINITY, -1}}, x_u[2][2] = {{INFINITY, 1}, {INFINITY, 1}};
    if (b200_device_count() <= 0) { printf("no CUDA device: %s\n", b200_last_error()); return 0; }

    b200qp_settings st;
    b200qp_set_default_settings_dense(&st);
    b200qp_handle* h = NULL;
    if (b200qp_setup_dense(&h, batch, n, p, m, &P[0][0], &c[0][0], &A[0][0], &b[0][0], &G[0][0], &h_l[0][0], &h_u[0][0], &x_l[0][0], &x_u[0][0],
                           &st, /*device=*/0, /*on_device=*/0) < 0) { printf("setup failed: %s\n", b200_last_error()); return 1; }
    if (b200qp_solve(h) < 0) { printf("solve failed: %s\n", b200_last_error()); return 1; }
    double x[2][2], y[2][1];
    b200qp_get_result(h, &x[0][0], &y[0][0], NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, 0);
    b200qp_info info[2];
    b200qp_get_info(h, info);
    for (int k = 0; k < batch; k++)
        printf("instance %d: status %d, %d iterations, x = (%.7f, %.7f), y = %.7f   [reference: (0.4285714, 0.2142857), -1.5714286]\n",
               k, info[k].status, (int)info[k].iter, x[k][0], x[k][1], y[k][0]);
    b200qp_cleanup(h);
    return 0;
}
