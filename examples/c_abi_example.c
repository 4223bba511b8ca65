/*
 * c_abi_example.c -- a plain C consumer of libebm_cuda.so (include/ebm_cuda.h): 64 classic members, 2 years, L0
 * diagnostics.  This is the call sequence the Julia extension performs with `ccall`.
 *
 *   gcc -std=c11 -Iinclude examples/c_abi_example.c -Lenergybalancemodel.jl_b200/lib -lebm_cuda -lm \
 *       -Wl,-rpath,$PWD/energybalancemodel.jl_b200/lib -o c_abi_example && ./c_abi_example
 *
 * Without a CUDA device the call fails with EBM_ERR_CUDA and prints the library's message (no CPU fallback).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "ebm_cuda.h"

int main(void) {
  enum { NX = 100, NT = 2000, DUR = 2, NMEM = 64 };
  static double x[NX], t[NT];
  for (int j = 0; j < NX; ++j) x[j] = (2.0 * j + 1.0) / (2.0 * NX);   /* SpaceTime{identity}: cell centres */
  for (int i = 0; i < NT; ++i) t[i] = (2.0 * i + 1.0) / (2.0 * NT);   /* st.t */
  ebm_grid_t grid = {NX, NT, DUR, 0, 522, 1548, x, t};                /* winter.inx / summer.inx at nt = 2000 */

  ebm_classic_params_t* par = malloc(sizeof(*par) * NMEM);
  ebm_forcing_t* forc = calloc(NMEM, sizeof(*forc));
  double* E0 = malloc(sizeof(double) * NMEM * NX);
  double* Tg0 = malloc(sizeof(double) * NMEM * NX);
  double* diag = malloc(sizeof(double) * NMEM * DUR * EBM_NSEASON * EBM_NDIAG);
  double* Ef = malloc(sizeof(double) * NMEM * NX);
  double* Tgf = malloc(sizeof(double) * NMEM * NX);
  for (int m = 0; m < NMEM; ++m) {
    ebm_classic_params_t p = {0.6, 193.0, 2.1, 9.8, 420.0, 338.0, 240.0, 0.7, 0.1, 0.4, 4.0, 2.0, 9.5, 0.098, 1e-5};
    par[m] = p;                                                       /* default_parameters(:Classic) */
    const double F = -10.0 + 20.0 * m / (NMEM - 1);
    forc[m].base = forc[m].peak = forc[m].cool = F;                   /* Forcing(F) */
    for (int j = 0; j < NX; ++j) { E0[m * NX + j] = 98.0; Tg0[m * NX + j] = 10.0; }
  }
  ebm_options_t opt = {-1, 1, 0, 0, 0, 0, 0.0, 0, 0, 0};
  ebm_classic_outputs_t out = {diag, NULL, NULL, Ef, Tgf, NULL};
  printf("%s, %d device(s)\n", ebm_version(), ebm_device_count());
  const int32_t rc = ebm_classic_run(&grid, NMEM, par, forc, E0, Tg0, &opt, &out);
  if (rc != EBM_OK) {
    printf("ebm_classic_run failed (%d): %s\n", rc, ebm_last_error());
    return rc == EBM_ERR_CUDA ? 0 : 1;   /* expected on a machine without a GPU */
  }
  for (int m = 0; m < NMEM; m += 21) {
    const double* d = diag + ((size_t)(m * DUR + (DUR - 1)) * EBM_NSEASON + EBM_SEASON_AVG) * EBM_NDIAG;
    printf("member %2d  F=%+6.2f  year-2 annual mean: T=%.3f  E=%.3f  ice area=%.3f  ice edge x=%.3f\n", m,
           forc[m].base, d[EBM_DIAG_MEAN_T], d[EBM_DIAG_MEAN_E], d[EBM_DIAG_ICE_AREA], d[EBM_DIAG_ICE_EDGE]);
  }
  ebm_shutdown();
  return 0;
}
