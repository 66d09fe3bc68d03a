/* field_solver.c -- CPU oracle of the Yee FDTD field update that sits between the J halo and the E/B
 * halo of a step.
 *
 * TEST INFRASTRUCTURE ONLY (see nix_oracle.h).  PARITY UNPINNED BY THE REFERENCE: amanotk/nix ships no
 * field solver -- Application::push() is an empty virtual (application.hpp:343-346) and the Maxwell
 * update lives in the downstream application.  What the reference DOES fix is the staggering of the six
 * components inside uf, through the way it colocates them for output
 * (xtensor_packer3d.hpp:279-302: Ex(ix) and Ex(ix+1) average to the cell centre, Bx(iy..iy+1, iz..iz+1)
 * average to it, ...) and through the half-grid gather of the push (node h at xmin + h*dx, DESIGN.md
 * section 2).  With cell centres at (i + 1/2) dx and edges at i dx that is
 *     Ex (c,c,e)  Ey (c,e,c)  Ez (e,c,c)      Bx (e,e,c)  By (e,c,e)  Bz (c,e,e)       (z,y,x)
 * and J is staggered like E (esirkepov.hpp:177-237: Jx(i) is the flux through the LOWER x face of cell
 * i, test_esirkepov.cpp:1018-1022).  The update below is the standard second-order leapfrog on that
 * lattice, written so that the GPU kernel can repeat every rounding (no contraction, fixed order):
 *     B -= c dt curl E      E += c dt curl B - cfj dt J
 * Pinned by known answers instead (tests/test_field_solver.py): discrete plane-wave eigenmodes with the
 * Yee dispersion relation, div B = 0 and Gauss's law div E = rho preserved to round-off together with
 * the Esirkepov deposit.
 */
#include "nix_oracle.h"

#include <stddef.h>

#define UF(z, y, x, c) uf[((((size_t)(z)) * my + (y)) * mx + (x)) * 6 + (c)]
#define UJ(z, y, x, c) uj[((((size_t)(z)) * my + (y)) * mx + (x)) * 4 + (c)]

/* B(iz,iy,ix) -= c dt curl E on cells [nb-ext, nb+N-1+ext] of every axis; reads E one cell below */
void nixo_fdtd_push_bfd(double* uf, const int* dims, int nb, const double* del, double cc, double delt, int ext)
{
  const int    my = dims[1] + 2 * nb, mx = dims[2] + 2 * nb;
  const double cz = cc * delt / del[0], cy = cc * delt / del[1], cx = cc * delt / del[2];
  for (int iz = nb - ext; iz <= nb + dims[0] - 1 + ext; iz++)
    for (int iy = nb - ext; iy <= nb + dims[1] - 1 + ext; iy++)
      for (int ix = nb - ext; ix <= nb + dims[2] - 1 + ext; ix++) {
        const double ex = UF(iz, iy, ix, 0), ey = UF(iz, iy, ix, 1), ez = UF(iz, iy, ix, 2);
        /* Bx: dEz/dy - dEy/dz    By: dEx/dz - dEz/dx    Bz: dEy/dx - dEx/dy */
        UF(iz, iy, ix, 3) = UF(iz, iy, ix, 3) - (cy * (ez - UF(iz, iy - 1, ix, 2)) - cz * (ey - UF(iz - 1, iy, ix, 1)));
        UF(iz, iy, ix, 4) = UF(iz, iy, ix, 4) - (cz * (ex - UF(iz - 1, iy, ix, 0)) - cx * (ez - UF(iz, iy, ix - 1, 2)));
        UF(iz, iy, ix, 5) = UF(iz, iy, ix, 5) - (cx * (ey - UF(iz, iy, ix - 1, 1)) - cy * (ex - UF(iz, iy - 1, ix, 0)));
      }
}

/* E(iz,iy,ix) += c dt curl B - cfj dt J on the interior cells; reads B one cell above */
void nixo_fdtd_push_efd(double* uf, const double* uj, const int* dims, int nb, const double* del, double cc,
                        double delt, double cfj)
{
  const int    my = dims[1] + 2 * nb, mx = dims[2] + 2 * nb;
  const double cz = cc * delt / del[0], cy = cc * delt / del[1], cx = cc * delt / del[2];
  const double cj = cfj * delt;
  for (int iz = nb; iz <= nb + dims[0] - 1; iz++)
    for (int iy = nb; iy <= nb + dims[1] - 1; iy++)
      for (int ix = nb; ix <= nb + dims[2] - 1; ix++) {
        const double bx = UF(iz, iy, ix, 3), by = UF(iz, iy, ix, 4), bz = UF(iz, iy, ix, 5);
        /* Ex: dBz/dy - dBy/dz    Ey: dBx/dz - dBz/dx    Ez: dBy/dx - dBx/dy */
        UF(iz, iy, ix, 0) = (UF(iz, iy, ix, 0) + (cy * (UF(iz, iy + 1, ix, 5) - bz) - cz * (UF(iz + 1, iy, ix, 4) - by))) - cj * UJ(iz, iy, ix, 1);
        UF(iz, iy, ix, 1) = (UF(iz, iy, ix, 1) + (cz * (UF(iz + 1, iy, ix, 3) - bx) - cx * (UF(iz, iy, ix + 1, 5) - bz))) - cj * UJ(iz, iy, ix, 2);
        UF(iz, iy, ix, 2) = (UF(iz, iy, ix, 2) + (cx * (UF(iz, iy, ix + 1, 4) - by) - cy * (UF(iz, iy + 1, ix, 3) - bx))) - cj * UJ(iz, iy, ix, 3);
      }
}

/* sum of E^2 and of B^2 over the interior cells, in row-major order (the GPU sums in another order:
 * compare to ~1e-13 relative) */
void nixo_fdtd_energy(const double* uf, const int* dims, int nb, double* e2b2)
{
  const int my = dims[1] + 2 * nb, mx = dims[2] + 2 * nb;
  double    e2 = 0.0, b2 = 0.0;
  for (int iz = nb; iz <= nb + dims[0] - 1; iz++)
    for (int iy = nb; iy <= nb + dims[1] - 1; iy++)
      for (int ix = nb; ix <= nb + dims[2] - 1; ix++) {
        for (int c = 0; c < 3; c++) e2 += UF(iz, iy, ix, c) * UF(iz, iy, ix, c);
        for (int c = 3; c < 6; c++) b2 += UF(iz, iy, ix, c) * UF(iz, iy, ix, c);
      }
  e2b2[0] = e2;
  e2b2[1] = b2;
}
