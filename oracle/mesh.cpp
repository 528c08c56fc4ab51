// TEST INFRASTRUCTURE ONLY (see fe_oracle.hpp).  Regional cube mesh of one tile.
#include "fe_oracle.hpp"

#include <limits>
#include <stdexcept>

namespace feo {

// mesh/scale_mesh_cubedom3d.F90:339-469 (setupLocalDom), :545-602 (coord_conv, calc_normal);
// mesh/scale_mesh_base3d.F90:175-331 (setGeometricInfo);
// mesh/scale_meshutil_3d.F90:38-118 (genCubeDomain), :264-475 (BuildInteriorMap: nearest-node
// matching between a face and the neighbour's face), :478-641 (genPatchBoundaryMap: halo slots).
void Mesh::init_cube(const Element& e, int nex, int ney, int nez, double xmin, double xmax, double ymin,
                     double ymax, double zmin, double zmax, const double* FZ, const bool per[3]) {
  NeX = nex; NeY = ney; NeZ = nez; Ne = nex * ney * nez; Ne2D = nex * ney;
  NeA = Ne + 2 * (nex + ney) * nez + 2 * nex * ney;
  for (int d = 0; d < 3; ++d) periodic[d] = per[d];
  const int Np = e.Np, np = e.np, Nfp = e.Nfp, NfpTot = e.NfpTot;
  // vertices (genCubeDomain)
  vec vx(nex + 1), vy(ney + 1), vz(nez + 1);
  for (int i = 0; i <= nex; ++i) vx[i] = (xmax - xmin) * double(i) / double(nex) + xmin;
  for (int j = 0; j <= ney; ++j) vy[j] = (ymax - ymin) * double(j) / double(ney) + ymin;
  for (int k = 0; k <= nez; ++k) vz[k] = FZ ? FZ[k] : (zmax - zmin) * double(k) / double(nez) + zmin;

  for (int d = 0; d < 3; ++d) pos[d].resize(size_t(Np) * Ne);
  E11.resize(size_t(Np) * Ne); E22 = E11; E33 = E11; J = E11;
  nx.assign(size_t(NfpTot) * Ne, 0.0); ny = nx; nz = nx; Fscale = nx;
  emap2d.resize(Ne);
  for (int kk = 0; kk < nez; ++kk) for (int jj = 0; jj < ney; ++jj) for (int ii = 0; ii < nex; ++ii) {
    int ke = ii + jj * nex + kk * nex * ney;
    emap2d[ke] = ii + jj * nex;
    double x0 = vx[ii], x1 = vx[ii + 1], y0 = vy[jj], y1 = vy[jj + 1], z0 = vz[kk], z1 = vz[kk + 1];
    double xX = 0.5 * (x1 - x0), yY = 0.5 * (y1 - y0), zZ = 0.5 * (z1 - z0);
    double Jac = xX * (yY * zZ);
    double e11 = (yY * zZ) / Jac, e22 = (xX * zZ) / Jac, e33 = (xX * yY) / Jac;
    for (int k = 0; k < np; ++k) for (int j = 0; j < np; ++j) for (int i = 0; i < np; ++i) {
      size_t n = size_t(i + j * np + k * np * np) + size_t(ke) * Np;
      pos[0][n] = x0 + 0.5 * (e.x1d[i] + 1.0) * (x1 - x0);
      pos[1][n] = y0 + 0.5 * (e.x1d[j] + 1.0) * (y1 - y0);
      pos[2][n] = z0 + 0.5 * (e.x1d[k] + 1.0) * (z1 - z0);
      E11[n] = e11; E22[n] = e22; E33[n] = e33; J[n] = Jac;
    }
    // normals: -E(2,:), +E(1,:), +E(2,:), -E(1,:), -E(3,:), +E(3,:), normalised; Fscale = sJ/J
    const double nv[6] = {-e22, e11, e22, -e11, -e33, e33};
    for (int f = 0; f < 6; ++f) {
      double sj = std::sqrt(nv[f] * nv[f]);
      for (int fp = 0; fp < Nfp; ++fp) {
        size_t n = size_t(f * Nfp + fp) + size_t(ke) * NfpTot;
        double nn = nv[f] / sj;
        if (f == 1 || f == 3) nx[n] = nn; else if (f == 0 || f == 2) ny[n] = nn; else nz[n] = nn;
        Fscale[n] = (sj * Jac) / Jac;
      }
    }
  }
  Gsqrt.assign(size_t(Np) * NeA, 1.0);
  G13.assign(size_t(Np) * NeA, 0.0); G23 = G13;
  GsqrtH.assign(size_t(Nfp) * Ne2D, 1.0);

  // --- interior maps by nearest-node search against the neighbour's facing face
  vmapM.resize(size_t(NfpTot) * Ne); vmapP.resize(size_t(NfpTot) * Ne);
  const int opp[6] = {2, 3, 0, 1, 5, 4};
  for (int ke = 0; ke < Ne; ++ke) {
    int ii = ke % nex, jj = (ke / nex) % ney, kk = ke / (nex * ney);
    int nb[6] = {jj > 0 ? ke - nex : -1, ii < nex - 1 ? ke + 1 : -1, jj < ney - 1 ? ke + nex : -1,
                 ii > 0 ? ke - 1 : -1, kk > 0 ? ke - nex * ney : -1, kk < nez - 1 ? ke + nex * ney : -1};
    for (int f = 0; f < 6; ++f) {
      int ke2 = nb[f] >= 0 ? nb[f] : ke, f2 = nb[f] >= 0 ? opp[f] : f;  // EToE/EToF default to self
      for (int fp = 0; fp < Nfp; ++fp) {
        int idM = e.Fmask[f * Nfp + fp] + ke * Np;
        vmapM[size_t(f * Nfp + fp) + size_t(ke) * NfpTot] = idM;
        double best = std::numeric_limits<double>::max(); int arg = -1;
        for (int fq = 0; fq < Nfp; ++fq) {
          int id2 = e.Fmask[f2 * Nfp + fq] + ke2 * Np;
          double dx = pos[0][idM] - pos[0][id2], dy = pos[1][idM] - pos[1][id2], dz = pos[2][idM] - pos[2][id2];
          double d2 = dx * dx + dy * dy + dz * dz;
          if (d2 < best) { best = d2; arg = id2; }   // minloc: first minimum
        }
        vmapP[size_t(f * Nfp + fp) + size_t(ke) * NfpTot] = arg;
      }
    }
  }
  // --- patch boundary: halo slots in order of tile face 1..6, elements ascending
  const int fsz[6] = {nex * nez, ney * nez, nex * nez, ney * nez, nex * ney, nex * ney};
  halo_off[0] = 0;
  for (int f = 0; f < 6; ++f) halo_off[f + 1] = halo_off[f] + fsz[f] * Nfp;
  Nhalo = halo_off[6];
  vmapB.resize(Nhalo);
  const double TOL = 1.0e-12;
  int counter = 0;
  for (int b = 0; b < 6; ++b) {
    for (int ke = 0; ke < Ne; ++ke) {
      // face-averaged coordinate compared with the tile bound (eval_domain_boundary)
      int d = (b == 0 || b == 2) ? 1 : (b == 1 || b == 3) ? 0 : 2;
      double bound = b == 0 ? ymin : b == 1 ? xmax : b == 2 ? ymax : b == 3 ? xmin : b == 4 ? vz[0] : vz[nez];
      double rnorm = d == 0 ? 1.0 / (xmax - xmin) : d == 1 ? 1.0 / (ymax - ymin) : 1.0 / std::fabs(vz[nez] - vz[0]);
      int f0 = b < 4 ? 0 : 4, f1 = b < 4 ? 4 : 6;
      for (int f = f0; f < f1; ++f) {
        double s = 0.0;
        for (int fp = 0; fp < Nfp; ++fp) s += pos[d][e.Fmask[f * Nfp + fp] + size_t(ke) * Np];
        s /= double(Nfp);
        if (std::fabs(s - bound) * rnorm < TOL) {
          for (int fp = 0; fp < Nfp; ++fp) {
            vmapP[size_t(f * Nfp + fp) + size_t(ke) * NfpTot] = Np * Ne + counter;
            vmapB[counter] = e.Fmask[f * Nfp + fp] + ke * Np;
            ++counter;
          }
        }
      }
    }
    if (counter != halo_off[b + 1]) throw std::runtime_error("halo bookkeeping mismatch");
  }
  // tile graph of a single tile (buildGlobalMap): periodic -> opposite face, else the same face
  for (int f = 0; f < 6; ++f) {
    bool pr = (f == 1 || f == 3) ? per[0] : (f == 0 || f == 2) ? per[1] : per[2];
    nbr_face[f] = pr ? opp[f] : f;
  }
}

// data/scale_meshfieldcomm_base.F90 (extract_bounddata -> same-rank copy in exchange_core -> set_bounddata)
void Mesh::exchange_halo(const Element& e, double* q) const {
  const size_t base = size_t(e.Np) * Ne;
  for (int f = 0; f < 6; ++f) {
    int fo = nbr_face[f];
    int n = halo_off[f + 1] - halo_off[f];
    for (int m = 0; m < n; ++m) q[base + halo_off[f] + m] = q[vmapB[halo_off[fo] + m]];
  }
}

}  // namespace feo
