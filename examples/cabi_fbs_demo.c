/*
 * libgigpower_b200 from plain C: no Python, no torch, host buffers only.
 *
 * The feeder tables (a GpFbsDesc, i.e. what gigpower_b200/flatten.py produces from the parsed .dss
 * file) and the scenario loads come from a generated header, see tests/test_cabi_c_demo.py which
 * writes it for the IEEE 13-bus feeder, compiles this file with gcc against include/gigpower_b200.h
 * and checks the printed voltages against the oracle.  Build by hand:
 *
 *   gcc -O2 -I include -I <dir of feeder_data.h> examples/cabi_fbs_demo.c \
 *       -L gigpower_b200 -lgigpower_b200 -Wl,-rpath,$PWD/gigpower_b200 -o cabi_fbs_demo
 *
 * Replaces, for a caller written in a compiled language, SolutionFBS.__init__ + solve()
 * (reference src/solution_fbs.py:27-136) for S scenarios at once.
 * Exit code: 0 ok, 3 library error (message on stderr) - e.g. no CUDA device: there is no CPU fallback.
 */
#include <stdio.h>
#include <stdlib.h>

#include "gigpower_b200.h"
#include "feeder_data.h"   /* GP_DEMO_* arrays and sizes */

static int die(const char* what, int rc) {
    char msg[512];
    gp_last_error(msg, sizeof msg);
    fprintf(stderr, "%s failed (%d): %s\n", what, rc, msg);
    return 3;
}

int main(void) {
    GpFbsDesc d;
    GpFeeder* f = NULL;
    const int64_t S = GP_DEMO_S, ld = GP_DEMO_S;
    int rc, s, r;

    d.n_steps = GP_DEMO_N_STEPS;
    d.n_vrows = GP_DEMO_N_VROWS;
    d.n_irows = GP_DEMO_N_IROWS;
    d.n_srows = GP_DEMO_N_SROWS;
    d.n_child = GP_DEMO_N_CHILD;
    d.n_lines = GP_DEMO_N_LINES;
    d.bus_vrow = gp_demo_bus_vrow;
    d.par_vrow = gp_demo_par_vrow;
    d.edge_irow = gp_demo_edge_irow;
    d.load_srow = gp_demo_load_srow;
    d.kind = gp_demo_kind;
    d.child_begin = gp_demo_child_begin;
    d.child_count = gp_demo_child_count;
    d.child_irow = gp_demo_child_irow;
    d.z = gp_demo_z;
    d.cappu = gp_demo_cappu;
    d.vr_gamma = gp_demo_vr_gamma;
    for (r = 0; r < 3; ++r) {
        d.root_vrow[r] = gp_demo_root_vrow[r];
        d.root_load_srow[r] = gp_demo_root_load_srow[r];
        d.root_cappu[r] = gp_demo_root_cappu[r];
        d.zip[r] = gp_demo_zip[r];
    }
    for (r = 0; r < 6; ++r) d.vref[r] = gp_demo_vref[r];
    d.line_irow = gp_demo_line_irow;
    d.line_tx_vrow = gp_demo_line_tx_vrow;
    d.line_rx_vrow = gp_demo_line_rx_vrow;

    printf("libgigpower_b200 version %d\n", gp_version());
    rc = gp_fbs_create(&d, 0, &f);
    if (rc) return die("gp_fbs_create", rc);

    {
        double* V = (double*)malloc(sizeof(double) * 2 * d.n_vrows * ld);
        double* I = (double*)malloc(sizeof(double) * 2 * (d.n_irows > 0 ? d.n_irows : 1) * ld);
        double* loss = (double*)malloc(sizeof(double) * 2 * S);
        double* diff = (double*)malloc(sizeof(double) * S);
        int32_t* iters = (int32_t*)malloc(sizeof(int32_t) * S);
        int32_t* status = (int32_t*)malloc(sizeof(int32_t) * S);
        /* tolerance = |Vref[1]| * 1e-6, maxiter = 100 (solution_fbs.py:37, solution.py:26) */
        rc = gp_fbs_solve_batch_host(f, S, ld, gp_demo_spu, V, I, NULL, loss, iters, status, diff, GP_DEMO_TOL, 100);
        if (rc) return die("gp_fbs_solve_batch_host", rc);
        for (s = 0; s < S; ++s) printf("scenario %d iterations %d status %d loss %.17g %.17g\n", s, iters[s], status[s], loss[2 * s], loss[2 * s + 1]);
        for (r = 0; r < d.n_vrows; ++r)
            for (s = 0; s < S; ++s) printf("V %d %d %.17g %.17g\n", r, s, V[2 * (r * ld + s)], V[2 * (r * ld + s) + 1]);
        free(V); free(I); free(loss); free(diff); free(iters); free(status);
    }
    rc = gp_feeder_destroy(f);
    if (rc) return die("gp_feeder_destroy", rc);
    return 0;
}
