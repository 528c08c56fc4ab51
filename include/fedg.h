/* fedg.h -- C ABI of the B200-native DG dynamics hot path (drop-in for FE-Project's
 * atm_dyn_dgm_nonhydro3d dynamics step).
 *
 * The reference has no C/FFI boundary: its plug-in seams are Fortran type-bound procedures and
 * procedure pointers.  Each entry point below names the reference interface it stands behind
 * (paths relative to FE-Project's FElib/src); INTEGRATION.md shows the ISO_C_BINDING interface
 * block a maintainer adds on the Fortran side.
 *
 * Conventions (same as the reference, so a Fortran caller passes its arrays unchanged):
 *   - all arrays are column-major with the node index fastest: field(Np, NeA);
 *   - index maps (VMapM, VMapP, VMapB, EMap3Dto2D) are 1-based;
 *   - real(RP) is double; -DSINGLE builds of the reference are not supported;
 *   - host pointers are only read/written during the call that receives them: the context copies
 *     what it needs to the device, the caller keeps ownership;
 *   - one context per (MPI rank, local mesh) <-> one GPU; calls on one context come from one thread;
 *   - every function returns FEDG_OK or an error code; fedg_last_error() gives the message.  The
 *     Fortran shim turns a non-zero status into LOG_ERROR + PRC_abort, which is how the reference
 *     reports errors.
 * There is no CPU fallback: without a CUDA device fedg_create() fails with FEDG_ERR_CUDA.
 */
#ifndef FEDG_H_
#define FEDG_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fedg_ctx fedg_ctx;

enum {
  FEDG_OK = 0,
  FEDG_ERR_ARG = 1,         /* bad argument / inconsistent sizes */
  FEDG_ERR_CUDA = 2,        /* CUDA runtime failure (message holds cudaGetErrorString) */
  FEDG_ERR_UNSUPPORTED = 3, /* configuration outside what the kernels implement */
  FEDG_ERR_STATE = 4,       /* call order violated (e.g. update before dyn_init) */
  FEDG_ERR_COMM = 5         /* NCCL failure */
};

/* boundary-condition ids, mesh/scale_mesh_bndinfo.F90:48-54 */
enum { FEDG_BND_NOSPEC = 0, FEDG_BND_PERIODIC = 1, FEDG_BND_SLIP = 2, FEDG_BND_NOSLIP = 3 };

/* Reference element + local mesh of one tile.
 * Replaces reading LocalMesh3D / ElementBase3D components inside the tendency routines:
 *   mesh/scale_localmesh_3d.F90:31-68, mesh/scale_localmesh_base.F90:31-75,
 *   element/scale_element_base.F90:111-145,
 *   element/scale_element_operation_tensorprod3D.F90.erb:66-80 (D1D, Lift_mat, IntrpMat_VPOrdM1). */
typedef struct fedg_mesh_desc {
  int polyorder;            /* PolyOrder_h == PolyOrder_v (TensorProd3D requirement, tensorprod3D.F90.erb:112); 7 (every equation set),
                             * 5, 3, 1 (NONHYDRO3D_HEVE; the element-operation factory covers P1..P15, tensorprod3D.F90:578-632) */
  int Ne, NeA, NeX, NeY, NeZ, Ne2D;
  int Nhalo;                /* size(VMapB) */
  /* reference element */
  const double* D1D;            /* (np,np)   elem1D%Dx1                                     */
  const double* Lift;           /* (Np,NfpTot) elem3D%Lift                                  */
  const double* VPOrdM1;        /* (np,np)   1D IntrpMat_VPOrdM1 (V * invV with last row 0) */
  const double* IntWeight_lgl;  /* (Np)                                                     */
  /* geometry */
  const double* Escale;     /* (Np,Ne,3,3)   */
  const double* Fscale;     /* (NfpTot,Ne)   */
  const double* normal_fn;  /* (NfpTot,Ne,3) */
  const double* J;          /* (Np,Ne)       */
  const double* Gsqrt;      /* (Np,NeA)      */
  const double* GI3;        /* (Np,NeA,2)    */
  const double* GsqrtH;     /* (Nfp_v,Ne2D)  */
  const double* zlev;       /* (Np,Ne)       */
  /* connectivity, 1-based */
  const int* VMapM;         /* (NfpTot,Ne) */
  const int* VMapP;         /* (NfpTot,Ne) */
  const int* VMapB;         /* (Nhalo)     */
  const int* EMap3Dto2D;    /* (Ne)        */
  /* tile graph seen from this tile, faces in the order y-, x+, y+, x-, z-, z+
   * (MeshCubeDom3D%tileID_globalMap / tileFaceID_globalMap / PRCRank_globalMap,
   *  mesh/scale_meshutil_3d.F90:750-877): rank owning the neighbour tile and the neighbour's
   *  face id (1..6).  A face with no neighbour points to the own rank with the same face id. */
  int nbr_rank[6];
  int nbr_face[6];
  int my_rank;
  /* velocity boundary condition per tile face (FEDG_BND_*), applied only where the face has no
   * neighbour (fluid_dyn_solver/scale_atm_dyn_dgm_bnd.F90:788-838) */
  int vel_bc[6];
  /* SCALE constants (scale_const), passed in: GRAV may be overridden by PARAM_CONST */
  double GRAV, Rdry, CPdry, CVdry, PRES00, OHM;
  /* Cubed-sphere panel tile (MeshCubedSphereDom3D, mesh/scale_mesh_cubedspheredom3d.F90:599-678); all NULL / 0 for a
   * regional mesh.  panelID 1..6; GIJ: lcmesh%GIJ (Nfp_v,Ne2D,2,2) contravariant horizontal metric; gam (Np,NeA);
   * pos2D: lcmesh2D%pos_en (Nfp_v,Ne2D,2) = central angles (alpha, beta).  Gsqrt, GsqrtH above carry the horizontal
   * Jacobian.  This build covers the shallow-atmosphere approximation without topography (gam = 1, GI3 = 0,
   * Gsqrt = GsqrtH on the 3D nodes); anything else is rejected with FEDG_ERR_UNSUPPORTED. */
  const double* GIJ;
  const double* gam;
  const double* pos2D;
  int panelID;
} fedg_mesh_desc;

const char* fedg_last_error(void);
int fedg_version(void);

/* Create / destroy the device-resident context for one local mesh. */
int fedg_create(const fedg_mesh_desc* desc, fedg_ctx** out);
void fedg_destroy(fedg_ctx* ctx);

/* AtmDynDGMDriver_nonhydro3d%Init: fluid_dyn_solver/scale_atm_dyn_dgm_driver_nonhydro3d.F90:355-594
 * eqs_type: "NONHYDRO3D_HEVE" | "NONHYDRO3D_HEVI" (p = 7; flat or terrain-following mesh: with Gsqrt / GI3 of the descriptor different
 * from 1 / 0 the vertical-implicit solver carries GsqrtV, G13, G23, scale_atm_dyn_dgm_nonhydro3d_rhot_hevi_common_2.F90:111-1328)
 * | "GLOBALNONHYDRO3D_HEVE" | "GLOBALNONHYDRO3D_HEVI"
 * (p = 7, cubed-sphere panel tile: fluid_dyn_solver/scale_atm_dyn_dgm_globalnonhydro3d_rhot_heve.F90:338-600,
 * scale_atm_dyn_dgm_globalnonhydro3d_rhot_hevi.F90:337-583, 873-1066, fluxes scale_atm_dyn_dgm_nonhydro3d_rhot_heve_numflux.F90:1543-1772,
 * scale_atm_dyn_dgm_nonhydro3d_rhot_hevi_numflux.F90:606-834); tinteg_type: a timeint_rk scheme name
 * (common/scale_timeint_rk_butcher_tab.F90:27-67).  filter_h1D / filter_v1D are the (np,np) matrices
 * MFilter_h1D and MFilter_v1D of Setup_ModalFilter (tensorprod3D.F90.erb:160-181, 460-505); pass
 * NULL when modalfilter_flag == 0. */
int fedg_dyn_init(fedg_ctx* ctx, const char* eqs_type, const char* tinteg_type, double dt,
                  int modalfilter_flag, const double* filter_h1D, const double* filter_v1D);

/* MeshField3D%local(n)%val of the five prognostic variables, (Np,NeA) each, host <-> device. */
int fedg_set_prog(fedg_ctx* ctx, const double* DDENS, const double* MOMX, const double* MOMY,
                  const double* MOMZ, const double* DRHOT);
int fedg_get_prog(fedg_ctx* ctx, double* DDENS, double* MOMX, double* MOMY, double* MOMZ, double* DRHOT);

/* AUX_VARS used by the step (driver_nonhydro3d.F90:683-690).  THERM_hyd may be NULL: it is then
 * computed as in atm_dyn_dgm_nonhydro3d_common_calc_RHOT_hyd (nonhydro3d_common.F90:584-619).
 * Interior values only are required: the halo part is filled by a halo exchange, as the model
 * does after reading a restart file (model mod_atmos_vars.F90:553-636). */
int fedg_set_aux(fedg_ctx* ctx, const double* DENS_hyd, const double* PRES_hyd, const double* THERM_hyd,
                 const double* Rtot, const double* CVtot, const double* CPtot);
/* AUXDYNVARS3D DPhydDx, DPhydDy (driver_nonhydro3d.F90:323-326, 1060-1095); NULL = zero. */
int fedg_set_phyd_hgrad(fedg_ctx* ctx, const double* DPhydDx, const double* DPhydDy);
/* update_phyd_hgrad (driver_nonhydro3d.F90:1060-1095 -> atm_dyn_dgm_nonhydro3d_common_calc_phyd_hgrad_lc, nonhydro3d_common.F90:624-777):
 * the same two fields computed ON THE DEVICE from the registered PRES_hyd (minus PRES_hyd_ref (Np,NeA), NULL = 0), with the halo values
 * the last background-field exchange left there (fedg_set_aux / fedg_comm_init across NCCL tiles, fedg_group_exchange_aux across linked
 * local meshes).  Needed whenever PRES_hyd varies horizontally (baroclinic-wave initial state). */
int fedg_update_phyd_hgrad(fedg_ctx* ctx, const double* PRES_hyd_ref);
/* Physics tendencies handed to the dynamics step, (Np,NeA) each: DENS_tp, MOMX_tp, MOMY_tp, MOMZ_tp, RHOT_tp, RHOH_p of
 * add_phy_tend (driver_nonhydro3d.F90:843-857, 1098-1178; ENTOT_CONSERVE_SCHEME_FLAG = .false. form).  They are added
 * to the explicit tendency of every stage inside the stage kernel.  NULL or all-zero arrays switch the term off. */
int fedg_set_phy_tend(fedg_ctx* ctx, const double* DENS_tp, const double* MOMX_tp, const double* MOMY_tp,
                      const double* MOMZ_tp, const double* RHOT_tp, const double* RHOH_p);
/* Coriolis parameter (Nfp_v, Ne2D); NULL = zero. */
int fedg_set_coriolis(fedg_ctx* ctx, const double* coriolis);

/* AtmDynDGMDriver_nonhydro3d%Update (driver_nonhydro3d.F90:614-963), nsteps times, state resident
 * on the device. */
int fedg_dyn_update(fedg_ctx* ctx, int nsteps);
/* Same, called the way the reference driver calls Update: prognostic fields live in host arrays (Np, NeA);
 * their (Np, Ne) interior is copied to the device, advanced nsteps and copied back.  The halo part [Ne+1:NeA] of the host
 * arrays is neither read nor written: the exchange of every stage rebuilds it on the device. */
int fedg_dyn_update_host(fedg_ctx* ctx, double* DDENS, double* MOMX, double* MOMY, double* MOMZ,
                         double* DRHOT, int nsteps);

/* Pipelined form of fedg_dyn_update_host for a caller that keeps up to three sets of host arrays (slot 0 / 1 / 2): the call returns once the
 * work is queued -- upload on a copy stream, nsteps on the compute stream, download on a second copy stream -- and
 * fedg_dyn_update_host_wait(slot) blocks until the outputs of that slot are complete.  While one slot downloads, the other uploads
 * and computes: PCIe runs in both directions at once.  Inputs are read and outputs written between the call and its wait; pinned
 * (page-locked) host memory is needed for the copies to be asynchronous.  In / out arrays may be the same. */
int fedg_dyn_update_host_async(fedg_ctx* ctx, const double* DDENS, const double* MOMX, const double* MOMY, const double* MOMZ,
                               const double* DRHOT, double* DDENS_out, double* MOMX_out, double* MOMY_out, double* MOMZ_out,
                               double* DRHOT_out, int nsteps, int slot);
int fedg_dyn_update_host_wait(fedg_ctx* ctx, int slot);

/* ---- stage-level seams: a driver that keeps the reference's own stage loop (driver_nonhydro3d.F90:703-921) over DEVICE-RESIDENT
 * buffers, no host copy per stage.  One step is
 *     fedg_rk_store_var0
 *     do stage = 1, nstage                                   (1-based, as tint%Advance takes it)
 *        [HEVI] fedg_cal_vi_dev(stage); fedg_rk_store_implicit(stage)
 *        fedg_halo_start; fedg_halo_wait                      (optional: fedg_cal_tend_ex_dev does the exchange when it was not done)
 *        fedg_cal_tend_ex_dev(stage)
 *        fedg_rk_advance(stage)
 *     fedg_modalfilter_apply
 * and gives the state fedg_dyn_update(ctx, 1) gives (the fused path applies the same operations inside fewer kernels).
 *   fedg_rk_store_var0      timeint_rk%StoreVar0            common/scale_timeint_rk.F90:624
 *   fedg_rk_store_implicit  timeint_rk%StoreImplicit        common/scale_timeint_rk.F90:2510 (rk_storeimpl_general2D): q += impl_fac * k_im;
 *                           the column kernel of fedg_cal_vi_dev has produced that state together with k_im, this call makes it current
 *   fedg_rk_advance         timeint_rk%Advance_varlist      common/scale_timeint_rk.F90:536 -> :1182 (low storage), :2201 (general / IMEX)
 *   fedg_cal_tend_ex_dev    cal_tend_ex into tint%tend_buf2D_ex(:,:,:,tintbuf_ind)   driver_nonhydro3d.F90:815-828
 *   fedg_cal_vi_dev         cal_vi into tint%tend_buf2D_im, impl_fac = tint%Get_implicit_diagfac(stage)   driver_nonhydro3d.F90:738-753
 *   fedg_halo_start / _wait MeshFieldComm_Exchange(do_wait=.false.) / MeshFieldComm_Get + ApplyBC_PROGVARS_lc
 *                           model_framework/scale_model_var_manager.F90:458-487, driver_nonhydro3d.F90:770-808; a tendency call between
 *                           start and wait processes the interior elements first (HIDE_MPI_COMM_FLAG, driver_nonhydro3d.F90:859-895)
 *   fedg_modalfilter_apply  atm_dyn_dgm_modalfilter_apply   driver_nonhydro3d.F90:940-951
 *   fedg_rk_get_tend        reads tint%tend_buf2D_ex / _im of a stage back to host arrays (Np,Ne) (diagnostics, tests) */
int fedg_rk_store_var0(fedg_ctx* ctx);
int fedg_rk_store_implicit(fedg_ctx* ctx, int stage);
int fedg_rk_advance(fedg_ctx* ctx, int stage);
int fedg_cal_tend_ex_dev(fedg_ctx* ctx, int stage);
int fedg_cal_vi_dev(fedg_ctx* ctx, int stage);
int fedg_halo_start(fedg_ctx* ctx);
int fedg_halo_wait(fedg_ctx* ctx);
int fedg_modalfilter_apply(fedg_ctx* ctx);
int fedg_rk_get_tend(fedg_ctx* ctx, int implicit, int stage, double* DENS_dt, double* MOMX_dt, double* MOMY_dt, double* MOMZ_dt,
                     double* RHOT_dt);

/* atm_dyn_nonhydro3d_cal_tend_ex seam (driver_nonhydro3d.F90:152-199, 815-828): explicit tendency of
 * the state currently on the device, after halo exchange, pressure and boundary conditions, i.e.
 * what the driver stores in tint%tend_buf2D_ex at one stage.  Outputs are host arrays (Np,Ne). */
int fedg_cal_tend_ex(fedg_ctx* ctx, double* DENS_dt, double* MOMX_dt, double* MOMY_dt, double* MOMZ_dt,
                     double* RHOT_dt);

/* atm_dyn_nonhydro3d_cal_vi seam (driver_nonhydro3d.F90:201-250, 738-753; rhot_hevi.F90:772-965): vertical-implicit
 * tendency of the state on the device, one Newton iteration about var0 (host arrays (Np,Ne), the tint%var0_2D of the
 * reference).  impl_fac = a_im(s,s)*dt; impl_fac == 0 evaluates the vertical operator explicitly.  Needs
 * fedg_dyn_init("NONHYDRO3D_HEVI", ...).  Outputs are host arrays (Np,Ne). */
int fedg_cal_vi(fedg_ctx* ctx, double impl_fac, const double* DDENS0, const double* MOMX0, const double* MOMY0,
                const double* MOMZ0, const double* DRHOT0, double* DENS_dt, double* MOMX_dt, double* MOMY_dt,
                double* MOMZ_dt, double* RHOT_dt);

/* atm_dyn_dgm_nonhydro3d_common_calc_pressure (nonhydro3d_common.F90:350-393): PRES, DPRES (Np,Ne) of
 * the state on the device. */
int fedg_get_pres(fedg_ctx* ctx, double* PRES, double* DPRES);

/* Halo exchange + boundary condition of the prognostic variables, exposed for conformance tests:
 * MeshFieldComm_Exchange (model_framework/scale_model_var_manager.F90:458-487) followed by
 * ApplyBC_PROGVARS_lc (scale_atm_dyn_dgm_bnd.F90:270-367).  After the call fedg_get_prog returns
 * arrays whose halo part [Np*Ne, Np*Ne+Nhalo) is filled. */
int fedg_exchange_halo(fedg_ctx* ctx, int apply_bc);

/* Conservation monitors: sum(IntWeight_lgl * J * Gsqrt * f) for f = DDENS, ENGT, ENGK, ENGI, ENGP
 * (file/scale_file_monitor_meshfield.F90:176-213; fields of model mod_atmos_vars_container.F90:1281-1357).
 * Local-tile sums; out[5]. */
int fedg_monitor(fedg_ctx* ctx, double* out);

/* timeint_rk tables (common/scale_timeint_rk_butcher_tab.F90): sizes by fedg_rk_info, then coefficients
 * row-major: a_ex,a_im (nstage*nstage), b_ex,b_im (nstage), sig,gam ((nstage+1)*nstage). */
int fedg_rk_info(const char* scheme, int* nstage, int* tend_buf_size, int* low_storage, int* imex);
int fedg_rk_coef(const char* scheme, double* a_ex, double* b_ex, double* a_im, double* b_im, double* sig,
                 double* gam);

/* ElementOperationBase3D conformance entry points (element/scale_element_operation_base.F90:33-193),
 * host arrays, nelem elements at once; names: "Dx","Dy","Dz","Lift","VFilterPM1","ModalFilter".
 * in: (Np,nelem) or (NfpTot,nelem) for Lift; out: (Np,nelem). */
int fedg_elem_op(fedg_ctx* ctx, const char* name, const double* in, double* out, int nelem);

/* ElementOperationBase3D%Div (element/scale_element_operation_base.F90:106-118, scale_element_operation_tensorprod3D.F90.erb:299-329):
 * vec_out(:,1:3) = Dx vec_in(:,1), Dy vec_in(:,2), Dz vec_in(:,3); vec_out(:,4) = Lift vec_in_lift.  Host arrays vec_in (Np,3,nelem),
 * vec_in_lift (NfpTot,nelem), vec_out (Np,4,nelem); the caller combines them with Escale / Gsqrt as the tendency routines do
 * (FElib/test/FE/element_operation_hexahedral/test_element_operation_hexahedral.f90:131-139). */
int fedg_elem_div(fedg_ctx* ctx, const double* vec_in, const double* vec_in_lift, double* vec_out, int nelem);

/* Timing of the last fedg_dyn_update call measured with CUDA events on the context's stream:
 * ms_total, and the summed duration of the stage kernels only. */
int fedg_last_timing(fedg_ctx* ctx, double* ms_total, double* ms_stage_kernels, long* n_launches);

/* Multi-GPU: NCCL communicator over the ranks of the tile graph.  The caller obtains the 128-byte
 * unique id on rank 0 (fedg_comm_unique_id), broadcasts it with its own transport (MPI_Bcast in the
 * reference, torch.distributed in the tests) and every rank calls fedg_comm_init.  Replaces
 * MeshFieldCommBase Put/Exchange/Get over MPI (data/scale_meshfieldcomm_base.F90:58-139). */
int fedg_comm_unique_id(void* id128);
int fedg_comm_init(fedg_ctx* ctx, const void* id128, int rank, int nranks);

/* ---- numerical diffusion (AtmDyn_Nonhydro3D_Numdiff, fluid_dyn_solver/scale_atm_dyn_dgm_nonhydro3d_numdiff.F90:119-376) --------
 * fedg_numdiff_init: PARAM_ATMOS_DYN_NUMDIFF (ND_LAPLACIAN_NUM, ND_COEF_h, ND_COEF_v); the step is TIME_DT of fedg_dyn_init, the
 * velocity BC ids are those of fedg_mesh_desc, therm_bc[6] gives the thermal BC id per tile face (1 = ADIABAT,
 * mesh/scale_mesh_bndinfo.F90; NULL = none).  apply_in_update != 0: applied after every step inside fedg_dyn_update, where
 * the model calls it (model/atm_nonhydro3d/src/atmos/mod_atmos_dyn.F90:343-349).
 * fedg_numdiff_apply: %Apply on the state on the device (THERM, MOMZ, MOMX, MOMY, DENS in this order).
 * With apply_in_update the diffusion runs inside fedg_dyn_update only: fedg_group_update (several local meshes) returns
 * FEDG_ERR_UNSUPPORTED for a mesh configured that way instead of silently stepping without it. */
int fedg_numdiff_init(fedg_ctx* ctx, int nd_laplacian_num, double nd_coef_h, double nd_coef_v, const int* therm_bc,
                      int apply_in_update);
int fedg_numdiff_apply(fedg_ctx* ctx);

/* ---- sponge layer (AtmDynSpongeLayer, fluid_dyn_solver/scale_atm_dyn_dgm_spongelayer.F90:55-220) -------------------------
 * PARAM_ATMOS_DYN_SPONGELAYER: SL_WDAMP_TAU (< 0: 10 TIME_DT), SL_WDAMP_HEIGHT, SL_WDAMP_LAYER (> 0 overrides the height by
 * the first node of that element layer), SL_HORIVELDAMP_FLAG.  The Rayleigh damping enters the explicit tendency of every
 * stage inside the stage kernel (driver_nonhydro3d.F90:830-841). */
int fedg_sponge_init(fedg_ctx* ctx, double sl_wdamp_tau, double sl_wdamp_height, int sl_wdamp_layer, int sl_horiveldamp_flag);
/* The same on a mesh with topography: the damping profile is a function of the COMPUTATIONAL height lmesh%pos_en(:,:,3)
 * (spongelayer.F90:111, 168-173), which fedg_mesh_desc does not carry (its zlev is the real height); pos_en3: host array (Np,Ne). */
int fedg_sponge_init_pos(fedg_ctx* ctx, double sl_wdamp_tau, double sl_wdamp_height, int sl_wdamp_layer, int sl_horiveldamp_flag,
                         const double* pos_en3);

/* ---- several local meshes on one device (LOCAL_MESH_NUM > 1; cubed-sphere panels) ------------------------------
 * fedg_link_halo: the halo of tile face `face` (1..6) of `ctx` is filled from the interior of `src`, another local mesh on
 * the same device -- the same-rank path of MeshFieldCommBase_exchange_core (data/scale_meshfieldcomm_base.F90:870-895).
 * src_index(n): 1-based interior index (into the (Np,Ne) part of src's fields) feeding each halo slot of the face, in
 * halo order; it encodes tileID_globalMap / tileFaceID_globalMap and, for a negative face id, revert_hori
 * (data/scale_meshfieldcomm_cubedspheredom3d.F90:492-540).  rot(4,n) (may be NULL): per halo node the matrix
 * [r00 r01; r10 r11] that turns the source's (MOMX, MOMY) into the receiver's components, i.e.
 * LonLat2CSVec(own panel, own face node) o CS2LonLatVec(source panel, source node)
 * (MeshFieldCommCubedSphereDom3D_exchange :226-420, common/scale_cubedsphere_coord_cnv.F90:150-236, 314-401). */
int fedg_link_halo(fedg_ctx* ctx, int face, fedg_ctx* src, const int* src_index, const double* rot);
/* The same link when the source mesh lives on ANOTHER rank (the reference ships panel edges between ranks with MPI_Isend / Irecv
 * tagged 10*tileID+faceID, data/scale_meshfieldcomm_base.F90:870-884; here one NCCL send/recv pair per linked face inside one
 * group per exchange).  The receiving rank calls fedg_link_halo_recv(ctx, face, peer_rank, msg_id, rot); the rank that owns
 * the source mesh calls fedg_link_halo_send(src_ctx, peer_rank, msg_id, src_index, n) with the same src_index the local form
 * takes.  msg_id is any number both sides agree on and that is unique per linked face (e.g. 6*panel + face of the receiver):
 * NCCL matches the messages of a rank pair in ascending msg_id.  The communicator is the one of the FIRST mesh passed to
 * fedg_group_update / fedg_group_exchange_halo (fedg_comm_init on it). */
int fedg_link_halo_recv(fedg_ctx* ctx, int face, int peer_rank, int msg_id, const double* rot);
int fedg_link_halo_send(fedg_ctx* src_ctx, int peer_rank, int msg_id, const int* src_index, int n);
/* MeshFieldComm_Exchange of the prognostic variables (+ DPRES) over the local meshes of a rank: remote panel edges, then
 * fedg_exchange_halo of every mesh. */
int fedg_group_exchange_halo(fedg_ctx** ctxs, int n, int apply_bc);
/* The same exchange for the background fields DENS_hyd, PRES_hyd, THERM_hyd of the linked faces (the AUX_VARS exchange the model does
 * after the restart file is read, model mod_atmos_vars.F90:553-636); call once after fedg_set_aux on every mesh of the group.  Without
 * it a linked face keeps the own-face values fedg_set_aux leaves in the halo. */
int fedg_group_exchange_aux(fedg_ctx** ctxs, int n);
/* AtmDynDGMDriver_nonhydro3d%Update over the local meshes of a rank (the `do n = 1, LOCAL_MESH_NUM` loops of
 * driver_nonhydro3d.F90:703-921): every stage piece runs on all meshes before the next one starts, so that linked halos
 * see the neighbours' stage state.  HEVI equation sets. */
int fedg_group_update(fedg_ctx** ctxs, int n, int nsteps);

/* ---- tracer advection with a prescribed mass flux --------------------------------------------
 * AtmDynDGMDriver_trcadv3d_update with ONLY_TRACERADV_FLAG = .true. (fluid_dyn_solver/scale_atm_dyn_dgm_driver_trcadv3d.F90:312-559;
 * kernels of scale_atm_dyn_dgm_trcadvect3d_heve.F90:149-777): the mass flux is the momentum of the state registered with
 * fedg_set_prog, DDENS_TRC = DDENS0_TRC = DDENS.  FCT coefficient + TMAR limiters unless disable_limiter, tracer modal filter
 * (1D matrices as in fedg_dyn_init) at the last stage; low-storage explicit schemes (Advance_trcvar).  Flat regional mesh, one
 * tile, p = 3 or 7.  Parity with the CPU restatement: tests/test_gpu_tracer.py.
 * fedg_trcadv_update: QTRC (Np, NeA) host array, its (Np, Ne) interior advanced in place by nsteps; RHOQ_tp may be NULL. */
int fedg_trcadv_init(fedg_ctx* ctx, const char* tinteg_type, double dt, int modalfilter_flag, const double* filter_h1D,
                     const double* filter_v1D, int disable_limiter);
int fedg_trcadv_update(fedg_ctx* ctx, double* QTRC, const double* RHOQ_tp, int nsteps);
/* The coupled mode (ONLY_TRACERADV_FLAG = .false., driver_trcadv3d.F90:426-539).  fedg_trcadv_couple(ctx, 1): from now on every dynamics
 * stage of fedg_dyn_update saves the mass flux and the Rusanov coefficient x density with the Butcher weights
 * (atm_dyn_dgm_trcadvect3d_save_massflux / _cal_alphdens_dyn, trcadvect3d_heve.F90:343-401, 460-551, called at
 * driver_nonhydro3d.F90:900-917), DDENS0_TRC / DDENS_TRC are the density at the start of the step and after the RK loop (:926-937; the
 * modal filter of the step then runs after the density was taken, as in the reference).  fedg_trcadv_update_coupled advances one
 * tracer by one step with the averages of the LAST dynamics step (scheme, step and limiter of fedg_trcadv_init; no TMAR in this mode) and
 * rescales it to the filtered density (:530-537).  Flat regional mesh, one tile.  Not included: the negative fixer of the model. */
int fedg_trcadv_couple(fedg_ctx* ctx, int on);
int fedg_trcadv_update_coupled(fedg_ctx* ctx, double* QTRC, const double* RHOQ_tp);

/* ---- sample/advect3d (BASELINE config 1) --------------------------------------------------------------
 * `sparsemat` in ELL storage as the reference holds it (common/scale_sparsemat.F90:33-55, 100-250):
 * val(M*col_size), colIdx(M*col_size), slot-major l = i + (k-1)*M (:172), colIdx 1-based. */
typedef struct fedg_sparsemat {
  int M, N, col_size;
  const double* val;
  const int* colIdx;
} fedg_sparsemat;

/* sparsemat_matmul, ELL (common/scale_sparsemat.F90:439-474, 554-634): c(:,j) = A b(:,j) for nvec right-hand sides;
 * b is (N,nvec), c is (M,nvec), host arrays. */
int fedg_sparsemat_matmul(const fedg_sparsemat* A, const double* b, double* c, int nvec);

/* `sparsemat` in either storage format of the reference type (common/scale_sparsemat.F90:33-55):
 *   storage_format_id = 1 (SPARSEMAT_STORAGE_TYPEID_CSR, :68): val(nnz), colIdx(nnz), rowPtr(rowPtrSize = M + 1), all 1-based;
 *   storage_format_id = 2 (SPARSEMAT_STORAGE_TYPEID_ELL, :69): val(M*col_size), colIdx(M*col_size) as in fedg_sparsemat, rowPtr unused. */
typedef struct fedg_sparsemat_any {
  int storage_format_id;
  int M, N, nnz, col_size, rowPtrSize;
  const double* val;
  const int* colIdx;
  const int* rowPtr;
} fedg_sparsemat_any;

/* The three generic products of the type, either storage (host arrays in, host arrays out):
 *   fedg_sparsemat_matmul1    sparsemat_matmul1   (:355-383; CSR kernel :439-474, ELL :554-585)   c(M) = A b(N)
 *   fedg_sparsemat_matmul1_2  sparsemat_matmul1_2 (:386-408; CSR :476-512, ELL :587-634)           c(M) = A (b1 .* b2)
 *   fedg_sparsemat_matmul2    sparsemat_matmul2   (:411-431; CSR :514-552, ELL :636-664)           c(NQ,M) = A b(NQ,N), Fortran order
 * The sums run in the reference's order (CSR: entries of a row ascending; ELL: slots ascending). */
int fedg_sparsemat_matmul1(const fedg_sparsemat_any* A, const double* b, double* c);
int fedg_sparsemat_matmul1_2(const fedg_sparsemat_any* A, const double* b1, const double* b2, double* c);
int fedg_sparsemat_matmul2(const fedg_sparsemat_any* A, const double* b, double* c, int NQ);

/* Set-up of the advection run on the mesh of ctx: the four operators the sample passes to
 * advect3d_kernel_cal_tend (sample/advect3d/mod_advect3d_kernel.f90:34-44) and the timeint_rk scheme and step of
 * sample/advect3d/test_advect3d.f90 (TINTEG_SCHEME_TYPE, TIME_DT). */
int fedg_advect3d_init(fedg_ctx* ctx, const char* tinteg_type, double dt, const fedg_sparsemat* Dx,
                       const fedg_sparsemat* Dy, const fedg_sparsemat* Dz, const fedg_sparsemat* Lift);
/* q, u, v, w: MeshField3D%local(n)%val (Np,NeA), host <-> device (halo part is filled by the exchange). */
int fedg_advect3d_set(fedg_ctx* ctx, const double* q, const double* u, const double* v, const double* w);
int fedg_advect3d_get(fedg_ctx* ctx, double* q);
/* advect3d_kernel_cal_tend of the state on the device after the halo exchange: dqdt (Np,Ne), host array. */
int fedg_advect3d_cal_tend(fedg_ctx* ctx, double* dqdt);
/* nsteps passes of the stage loop of test_advect3d.f90:81-126 (exchange -> cal_tend -> Advance per stage),
 * state resident on the device; one step is captured once as a CUDA graph and replayed. */
int fedg_advect3d_update(fedg_ctx* ctx, int nsteps);

#ifdef __cplusplus
}
#endif
#endif /* FEDG_H_ */
