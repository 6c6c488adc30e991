// Modal-source eigensolve (see mode.cu)
#pragma once
int mode_solve(const double* eps_line, int n, double omega, double dl, int pol, double L0, double neff, int order,
               int averaged, double* vals, double* vecs);
