/*
 * gigpower_b200 -- C ABI of the B200-native batched power-flow engine.
 *
 * The reference (michael-sankur/gigpower) is pure Python and has no FFI of its
 * own on this path; the interface each entry point replaces is the Python
 * method cited beside it (paths relative to gigpower/src/gigpower/).  The
 * Python drop-in classes in gigpower_b200/ bind this header with ctypes; see
 * INTEGRATION.md for the stub a gigpower maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, a negative GP_E* code on failure;
 *     gp_last_error() gives the message of the calling thread's last failure;
 *   - no exceptions cross the ABI, no allocation of caller-visible buffers
 *     inside: the caller owns every data buffer;
 *   - "dev" pointers are CUDA device pointers on the feeder's device, "host"
 *     pointers are ordinary (preferably pinned) host memory;
 *   - complex values are interleaved (re, im) doubles; batched arrays are
 *     ROW-major [row][ld] with the scenario index innermost (ld >= S is the
 *     row pitch in scenarios), so that a warp reads 32 consecutive scenarios
 *     of one row with 128-bit loads;
 *   - a handle is not thread-safe; different handles are independent;
 *   - per-scenario failure is reported in status[] (GP_ST_*), never as a call
 *     failure.
 */
#ifndef GIGPOWER_B200_H
#define GIGPOWER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GP_OK 0
#define GP_EINVAL (-1)   /* bad argument / inconsistent descriptor */
#define GP_ECUDA (-2)    /* CUDA runtime error (message has the detail) */
#define GP_ENOMEM (-3)   /* host or device allocation failed */
#define GP_EJIT (-4)     /* the feeder-specialised kernel could not be generated, compiled or loaded */

#define GP_ST_CONVERGED 0
#define GP_ST_MAXITER 1
#define GP_ST_NONFINITE 2

/* edge kinds of an FBS step (solution_fbs.py:162-184 / :138-160) */
#define GP_EDGE_LINE 0   /* Line or Transformer (SyntheticLine with Z = 0, line.py:79-97) */
#define GP_EDGE_VR 1     /* one or more VoltageRegulators between the same two buses */

typedef struct GpFeeder GpFeeder;

/*
 * Flattened feeder for Forward-Backward Sweep.
 *
 * Replaces the object graph SolutionFBS walks: topo_order / adj / reverse_adj
 * (solution_fbs.py:29-33, 39-74; utils.py:130-181), Line.FZpu (line.py:38-46),
 * the phase matrices (circuit_element.py:47-51) and the cached load parameters
 * (solution.py:99-106).  Rows are PACKED: one V row per existing (bus, phase),
 * one I row per existing (line|transformer, phase), one spu row per (bus,
 * phase) that carries a load.
 *
 * Step k (0 <= k < n_steps) is the non-root bus at position k+1 of topo_order
 * together with the edge from its parent.  The forward sweep runs the steps
 * in order, the backward sweep in reverse order.
 */
typedef struct {
    int32_t n_steps;           /* nb - 1                                              */
    int32_t n_vrows;           /* P_b: rows of V                                      */
    int32_t n_irows;           /* P_l: rows of I                                      */
    int32_t n_srows;           /* P_L: rows of spu                                    */
    int32_t n_child;           /* entries of child_irow / 3                           */
    int32_t n_lines;           /* real lines (for Stx/Srx/loss, solution.py:222-230)  */
    /* per step, [n_steps][3]: row index or -1 when the phase does not exist */
    const int32_t* bus_vrow;   /* V rows of the step's bus                            */
    const int32_t* par_vrow;   /* V rows of its parent                                */
    const int32_t* edge_irow;  /* I rows of the incoming edge (all -1 for GP_EDGE_VR) */
    const int32_t* load_srow;  /* spu rows of the bus                                 */
    const int32_t* kind;       /* [n_steps] GP_EDGE_*                                 */
    const int32_t* child_begin;/* [n_steps] first child edge of the bus in child_irow */
    const int32_t* child_count;/* [n_steps] number of non-VR child edges              */
    const int32_t* child_irow; /* [n_child][3] I rows of those child edges, adj order */
    const double* z;           /* [n_steps][3][3][2] FZpu of the incoming edge        */
    const double* cappu;       /* [n_steps][3] capacitor pu of the bus                */
    const double* vr_gamma;    /* [n_steps][3] gamma per regulated phase, 0 elsewhere */
    /* root bus (topo_order[0]) */
    int32_t root_vrow[3];
    int32_t root_load_srow[3];
    double root_cappu[3];
    double vref[6];            /* solution_fbs.py:34-35, interleaved                  */
    double zip[3];             /* aPQ_p, aI_p, aZ_p (solution.py:44-45)               */
    /* real lines, [n_lines][3] each, -1 where the phase does not exist */
    const int32_t* line_irow;
    const int32_t* line_tx_vrow;
    const int32_t* line_rx_vrow;
} GpFbsDesc;

/* Library version, e.g. 100 = 0.1.0 */
int gp_version(void);

/* Copy the calling thread's last error message into buf (NUL-terminated). */
int gp_last_error(char* buf, size_t len);

/* Number of kernels this library has launched in this process (bench evidence). */
int64_t gp_launch_count(void);

/*
 * Process-wide switches.  "fbs_kernel" selects the FBS solve kernel (identical results):
 * 0 = automatic: the feeder-specialised on-chip kernel once gp_fbs_specialise has built it, else
 * the two-sweep streaming kernel (default); 1 = fused depth-first walk; 2 = two-sweep with three
 * phase lanes per scenario; 3 = two-sweep streaming (state in HBM); 4 = table-driven on-chip
 * kernel (state in shared memory; the solve fails with GP_EINVAL if not even one warp fits);
 * 5 = the feeder-specialised on-chip kernel (GP_EINVAL unless gp_fbs_specialise succeeded).
 * "nr3_kernel": 0 = automatic (default), 1 = one thread per scenario, 2 = three phase lanes per scenario
 * (Line / Transformer feeders only; automatic for small batches of mostly three-phase feeders).
 * "fbs_prefetch" / "nr3_prefetch": prefetch distance in tree steps of the streaming FBS kernel (default -1 =
 * automatic = off: every distance measured slower with the final kernel; 1..64 into L2, 100 + d into L1) and of
 * the NR3 kernels (default -1 = automatic: 2 for batches up to 65,536).
 * "nr3_ring": the phase-lane NR3 kernel stages the rows of the next two steps whose addresses do not depend on the
 *   sweep in flight (backward: the bus's X rows and load; forward: its front, r1 and X rows) in shared memory with
 *   cp.async, one 15 kB ring per warp: -1 (default) = on for feeders on which at least 5 of 6 buses are three-phase,
 *   0 = direct loads with an L2 prefetch "nr3_prefetch" steps ahead, 1 = always.
 * "fbs_share_rows": 1 (default) = in the streaming FBS kernel an I row that only repeats a child's current (a
 *   bus-phase without load and capacitor fed by one child edge) or stays zero makes no HBM round trips during
 *   the sweeps and is filled in once at the end; 0 = every row is stored and re-read each sweep.  Same results.
 * "fbs_kernel" 6 (on request, see "fbs_auto_walk"): the tree-walk kernel - the depth-first
 *   order below with the state of the bus the walk stands on held in registers and next-event operands requested
 *   ahead; "fbs_walk_blocks" (4, 5, 6, 8) = resident 128-thread blocks per SM it is compiled for (128 / 96 / 80 /
 *   64 registers per thread).  Per-scenario extras (taps, warm start, wpu) run on the two-sweep kernel.
 *   "fbs_auto_walk": 0 (default) = the automatic choice never takes the tree-walk kernel (it measured slower than the
 *   two-sweep kernel on B200); 1 = it does for feeders the walk covers.
 * "fbs_order": 1 (default) = the streaming FBS kernel runs a bus's forward step when a depth-first walk of the
 *   tree descends its edge and its backward step when the walk comes back up (children latest in topo order
 *   first); 0 = two full sweeps as the reference writes them (solution_fbs.py:92-121).  Same steps, same rows,
 *   same values - a forward step only ever sees its parent's forward voltage because of a bus's children only
 *   the earliest in topo order rewrites a phase of the parent, and that child is visited last - but a bus's
 *   voltage is re-read a few steps after it was written (an L2 hit) instead of a whole sweep later (DRAM).
 * "host_warps": warps per SM of the one-launch host path's kernel build (1..8, default 4; read when that build is
 *   made, i.e. at the first gp_fbs_solve_batch_host call of a handle).  The host path is PCIe-bound and a smaller
 *   wave of warps shortens the part of a call in which only one PCIe direction is busy.
 * "host_chunks": number of pipeline chunks of gp_fbs_solve_batch_host (0 = automatic).
 * "host_zero_copy": 1 (default) = with pinned (device-mapped) host buffers and a specialised feeder,
 *   gp_fbs_solve_batch_host is ONE launch that reads and writes the host buffers itself; 0 = always copy - solve - copy.
 * "jit_dfs": depth-first build of the feeder-specialised kernel - the tree is walked in the "fbs_order" order as nested
 *   straight-line code, a bus's forward voltage and accumulated current travel in registers between its descend
 *   and ascend, and the V rows live in the caller's V array (L2) instead of shared memory, which then holds only
 *   the I slots and a small stack: -1 (default) = automatic, when that at least doubles the warps per SM of the
 *   two-sweep build (34- and 37-bus class; not the 13-bus class); 0 = never; 1 = whenever the feeder has at
 *   most 64 steps.  Read by gp_fbs_create.  gp_fbs_solve_batch_host on such a handle runs a two-sweep build in its
 *   one-launch path (V rows must not cross PCIe once per sweep).
 * "host_read_ahead": 1 (default) = the one-launch host path uses its own build of the specialised kernel
 *   in which a warp issues the PCIe reads of its next tile before the stores of the tile it has solved
 *   (upload and download then overlap instead of alternating); 0 = the device-path kernel.
 */
int gp_set_option(const char* name, int value);

/*
 * Which kernel gp_fbs_solve_batch launches for this feeder under the current "fbs_kernel"
 * option (1..5 as above; bench.py uses it to pick the roofline that bounds the kernel), and
 * how many 32-scenario warps of the on-chip kernel fit in one SM (0 = none).
 */
int gp_fbs_active_kernel(GpFeeder* f);
int gp_fbs_onchip_warps(GpFeeder* f);

/*
 * Feeder-specialised kernel.  A flattened feeder fixes the whole sweep schedule, so for a
 * feeder that fits the on-chip kernel the library can write that schedule out as straight-line
 * CUDA C++ (row offsets as immediates, impedances as literals, absent phases and load-free
 * bus-phases elided, a scenario's loads held in registers), compile it for sm_100a with NVRTC
 * (dlopen'ed: no link-time dependency) and run it in place of the table-driven kernel.  Same
 * arithmetic, same outputs.
 *   gp_fbs_specialise     generate + compile + load for this handle (needs the device); GP_EJIT
 *                         with a message if NVRTC or the driver is unavailable - the handle then
 *                         keeps using the table-driven kernels
 *   gp_fbs_jit_source     host only: the CUDA source for a descriptor and `warps` warps per
 *                         block, with the I rows in shared memory (i_global = 0) or in global
 *                         memory (1); + 2 = the build with per-scenario regulator ratios and warm starts
 *                         that gp_fbs_solve_batch_ex launches on a specialised feeder; + 4 = the
 *                         depth-first build ("jit_dfs"; GP_EJIT when combined with 1 or 2 or asked of a
 *                         feeder with more than 64 steps);
 *                         returns its length (excluding the NUL) or a negative GP_E*;
 *                         copies at most cap-1 characters into buf when buf is not NULL
 *   gp_fbs_jit_warps      warps per block gp_fbs_specialise uses for this handle (0 = the feeder
 *                         does not fit): the V rows of a warp's 32 scenarios take shared memory, the
 *                         I rows too unless keeping them in the (L2-resident) I array lets more
 *                         warps run ("jit_i_global" option: -1 automatic, 0, 1)
 *   gp_jit_compile        host only: NVRTC-compile a source for sm_100a; the cubin size goes
 *                         to *cubin_bytes, the cubin itself into cubin (cap bytes) when not NULL
 */
int gp_fbs_specialise(GpFeeder* f);
/*
 * Host only (no device): the schedule the kernels are given for a descriptor.  events_out[2 e], [2 e + 1] =
 * {step, (level << 4) | flags} of the depth-first walk (flag bit 0: 1 = backward step on the way up, 0 =
 * forward step on the way down); kind_out[step] = the step's edge kind (bits 0..7) with the library's flag bits:
 * 8..15 phase mask of the mask-specialised body, 19..21 phases whose parent-voltage rewrite is dead
 * (solution_fbs.py:207-232: a sibling earlier in topo order overwrites it before anything reads it).  Either
 * pointer may be NULL; events_out needs 4 * n_steps entries, kind_out n_steps.  Returns the number of events
 * (2 * n_steps; 0 when the step table is not a tree in topo order) or a negative GP_E*.  The CPU tests run the
 * kernel's loop over these tables in NumPy and hold it to the oracle.
 */
int gp_fbs_schedule(const GpFbsDesc* desc, int32_t* events_out, int32_t* kind_out);
/*
 * Host only: the event table of the tree-walk kernel ("fbs_kernel" 6) for a descriptor, 20 int32 per event:
 * step, flags, e_row[3], ld_i[3], ld_v[3], st_ca[3], ld_s[3], pad[3] (struct WalkEv in csrc/gp_fbs.cu, which
 * documents the fields).  cap_events = room in events_out, in events.  Returns the number of events (between
 * n_steps and 2 n_steps; 0 = the feeder is not walked and runs on the two-sweep kernel) or a negative GP_E*.
 */
int gp_fbs_walk_events(const GpFbsDesc* desc, int32_t* events_out, int64_t cap_events);
int64_t gp_fbs_jit_source(const GpFbsDesc* desc, int32_t warps, int32_t i_global, char* buf, int64_t cap);
int gp_fbs_jit_warps(GpFeeder* f);
int gp_jit_compile(const char* source, void* cubin, int64_t cap, int64_t* cubin_bytes);

/*
 * Measured FP64 FMA throughput of the device in TFLOP/s (2 flop per DFMA, 8 independent
 * chains per thread, best of 5 launches timed with CUDA events): the denominator of the
 * FP64 roofline of the on-chip FBS kernel and of the NR3 elimination.
 */
int gp_measure_fp64_peak(int device, double* tflops);
/*
 * FP64 tensor-core probe (the "DMMA question" of DESIGN.md): out7[0] = TFLOP/s of mma.sync.m8n8k4.f64 alone (8
 * independent chains per warp), [1] = of DFMA alone, [2], [3] = of each while both are interleaved in one instruction
 * stream, [4..6] = the three kernel durations in ms.  Separate pipes would give t_both ~ max(t_mma, t_fma).
 */
int gp_measure_dmma(int device, double* out7);

/*
 * Build the device-resident feeder tables.  Replaces SolutionFBS.__init__'s
 * topology setup (solution_fbs.py:27-37).  `device` is a CUDA ordinal.
 */
int gp_fbs_create(const GpFbsDesc* desc, int device, GpFeeder** out);
int gp_feeder_destroy(GpFeeder* f);

/*
 * Solve S scenarios from flat start with Forward-Backward Sweep; replaces
 * SolutionFBS.solve() (solution_fbs.py:76-128) for every scenario at once.
 * Each scenario stops at the iteration at which a separate reference object
 * would (root-voltage mismatch <= tol, or maxiter).
 *
 *   spu_dev   [n_srows][ld] complex   per-scenario bus loads in pu (Solution.spu)
 *   V_dev     [n_vrows][ld] complex   out: bus voltages
 *   I_dev     [n_irows][ld] complex   out: line / transformer currents
 *   iters_dev [S] int32               out: iterations
 *   status_dev[S] int32               out: GP_ST_*
 *   diff_dev  [S] double              out: convergence_diff (solution_fbs.py:126)
 *
 * Asynchronous on `stream` (a cudaStream_t, NULL = default stream).
 */
int gp_fbs_solve_batch(GpFeeder* f, int64_t S, int64_t ld, const double* spu_dev,
                       double* V_dev, double* I_dev, int32_t* iters_dev,
                       int32_t* status_dev, double* diff_dev, double tol,
                       int32_t maxiter, void* stream);

/*
 * gp_fbs_solve_batch followed by the |V| and loss part of gp_fbs_post (either output may be
 * NULL) as ONE call: the feeder-specialised kernel writes both while the state is still on chip
 * (one launch); the other kernels run the solve and then the post kernels.  Same results.
 *   Vmag_dev [n_vrows][ld] double, loss_dev [S] complex
 */
int gp_fbs_solve_post_batch(GpFeeder* f, int64_t S, int64_t ld, const double* spu_dev, double* V_dev,
                            double* I_dev, int32_t* iters_dev, int32_t* status_dev, double* diff_dev,
                            double* Vmag_dev, double* loss_dev, double tol, int32_t maxiter, void* stream);

/*
 * Derived results of solved scenarios; replaces the tail of SolutionFBS.solve()
 * (solution_fbs.py:130-135; solution.py:177-233).  Any output pointer may be NULL.
 *
 *   sV_dev   [n_vrows][ld] complex    calc_sV
 *   Vmag_dev [n_vrows][ld] double     calc_Vmag
 *   Stx_dev  [n_irows][ld] complex    calc_Stx  (rows of real lines only)
 *   Srx_dev  [n_irows][ld] complex    calc_Srx
 *   loss_dev [S] complex              sum over real lines and phases of Stx - Srx
 */
int gp_fbs_post(GpFeeder* f, int64_t S, int64_t ld, const double* spu_dev,
                const double* V_dev, const double* I_dev, double* sV_dev,
                double* Vmag_dev, double* Stx_dev, double* Srx_dev,
                double* loss_dev, void* stream);

/*
 * Host-buffer entry point (what a gigpower caller without device code binds):
 * copies spu to the device in chunks, solves, and copies the requested results
 * back, overlapping copies and kernels on internal streams.  Host arrays use
 * the same [row][ld] layout; any of V/I/Vmag/loss may be NULL.  Synchronous.
 */
int gp_fbs_solve_batch_host(GpFeeder* f, int64_t S, int64_t ld, const double* spu_host,
                            double* V_host, double* I_host, double* Vmag_host,
                            double* loss_host, int32_t* iters_host, int32_t* status_host,
                            double* diff_host, double tol, int32_t maxiter);

/*
 * Optional per-scenario inputs of a solve.  Zero-initialise; a NULL
 * pointer / 0 keeps the behaviour of the plain entry points.
 *
 * gamma_dev: regulator ratios as a per-scenario input.  The reference fixes a regulator's ratio
 *   when the Solution object is built (voltage_regulator.py:10-39: gamma = 0.9 + 0.2 (tap + 16) / 32
 *   from RegControls.TapNumber()), so scenario s with ratios gamma[.][s] is "a fresh reference
 *   object built with those taps".  One row per regulated (step, phase), i.e. per non-zero entry of
 *   GpFbsDesc.vr_gamma / GpNr3Desc.vr_gamma, numbered in step order and, inside a step, phase
 *   order (gp_gamma_rows() = their count); [n_gamma_rows][ld] doubles.  FBS forward sweep
 *   V[reg] = gamma V[tx] (solution_fbs.py:138-160), backward sweep V[tx] = (1 / gamma) V[reg]
 *   (:186-205); NR3 ratio rows -gamma A_in + A_out = 0 (solution_nr3.py:480-483).
 *
 * warm_start: 1 = continue from an earlier solution instead of the flat start.  FBS: V_dev and I_dev
 *   hold that solution (what a reference object does when Vtest and iterations are reset and
 *   solve() is called again with new loads: V and I are not re-initialised, solution_fbs.py:76-90);
 *   NR3: X_dev holds it (SolutionNR3.solve() starts from self.XNR, solution_nr3.py:598).  The stop
 *   rules are unchanged, so a warm-started scenario usually needs fewer iterations and its result
 *   differs from the flat-started one within the solver tolerance.
 *
 * wpu_dev (FBS): the controllable reactive injection of calc_sV, sV = spu (aPQ + aI |V| + aZ |V|^2)
 *   - j cappu |V|^2 + j wpu (solution.py:203), per scenario and V row.  The reference keeps Solution.wpu at
 *   zero (circuit.py:164-171) but it is a plain attribute of the object; a caller that sets it - a volt-var
 *   droop evaluated between solves, volt_var_controller.py:20-38 - gets this term.  Runs on the streaming
 *   kernel; gp_fbs_post_ex gives the matching sV.  (gamma_dev and warm_start run on the feeder-specialised kernel
 *   once gp_fbs_specialise has built it - its own build, compiled at the first such call - and on the streaming
 *   kernel otherwise; wpu_dev and vv_* always take the streaming kernel.)
 * vv_ctl_dev, vv_points, vv_q (FBS): volt-var droop INSIDE the iteration.  On the flagged V rows
 *   Solution.wpu is not an input but wpu = VoltVARController.get_Q(|V|) (volt_var_controller.py:20-38: maxQ below
 *   V1, a ramp to 0 at V2, 0 up to V3, a ramp to minQ at V4, minQ above, sign flipped), re-evaluated from the bus's
 *   current voltage every time its power is formed - in every backward step of every sweep and in the final sV of
 *   gp_fbs_post_ex - which is what the reference's SolutionFBS does when its own controller objects are consulted
 *   before each calc_sV (oracle/make_golden_vvc_iter.py).  Runs on the streaming kernel.
 * wpu_dev (NR3): the same rows in SolutionNR3.  There wpu reaches the solve through change_KCL_matrices, which
 *   subtracts it from the constant term of BOTH KCL rows of a bus-phase (b = -load beta_S + cap gamma_S - wpu for
 *   the real and for the imaginary row alike, solution_nr3.py:559-588), and reaches sV as + j wpu
 *   (nr3_lib/map_output.py:60; gp_nr3_map_output_ex).  Runs on the thread-per-scenario kernel.
 *
 * n_peers, peer_vmag, peer_loss: the path's one exchange (the all-gather of per-scenario
 *   voltage magnitudes and losses across the GPUs of a box) fused into the solve.  Each destination is
 *   this rank's block inside the gather buffer of one GPU - its own or a peer's, mapped into this
 *   process with gp_peer_open - and receives what Vmag_dev / loss_dev receive, written by the same kernel
 *   epilogue with plain stores that travel over NVLink / NVSwitch while the next tile is being solved
 *   (no collective kernel, no SM taken from the solve).  With the feeder-specialised kernel that is one
 *   launch for solve + |V| + losses + all-gather; with the other kernels the |V| and loss kernels
 *   (gp_fbs_post / gp_nr3_map_output) do the stores.  Completion is the completion of the stream's
 *   work on every rank: consumers synchronise their stream and then cross a rank barrier.
 */
#define GP_MAX_PEERS 16
typedef struct {
    const double* gamma_dev;
    int32_t warm_start;
    int32_t n_peers;            /* fused all-gather of |V| and losses: number of destinations (0 = none)     */
    double* const* peer_vmag;   /* host array [n_peers] of device pointers, each a [n_vrows][ld] double block */
    double* const* peer_loss;   /* host array [n_peers] of device pointers, each a [S] complex block          */
    const double* wpu_dev;      /* [n_vrows][ld] doubles, Solution.wpu per scenario (NULL = zeros)            */
    const int32_t* vv_ctl_dev;  /* FBS: [n_vrows] int32 on the device, non-zero = a volt-var controller on this  */
                                /* V row (NULL = none); its wpu is get_Q(|V|) at every calc_sV, see above       */
    double vv_points[4];        /* the controllers' voltage breakpoints V1 < V2 <= V3 < V4 (pu)                  */
    double vv_q[2];             /* minQ, maxQ (pu)                                                              */
} GpSolveExtras;

/*
 * Peer memory for the fused all-gather (one process per GPU: CUDA IPC).
 *   gp_peer_alloc  device memory that other processes of the box can map; `handle_out` receives the
 *                  GP_PEER_HANDLE_BYTES opaque bytes to send them (any channel: torch.distributed
 *                  all_gather_object, a file, a pipe);
 *   gp_peer_open   map a peer's allocation into this process (on this process' current device `device`,
 *                  peer access over NVLink is enabled on demand); gp_peer_close unmaps it;
 *   gp_peer_push   copy-engine all-gather step for results that already sit in device memory: copies
 *                  `bytes` from src_dev to dst[q] for every q, asynchronously with respect to the host and ordered
 *                  in `stream` (DMA over NVLink, no SM involved).  The destinations are spread over
 *                  "peer_push_streams" internal copy streams (gp_set_option, 1..8, default 1: more were measured slower) that fork from and
 *                  join back into `stream`, so that several copy engines work at once.
 */
#define GP_PEER_HANDLE_BYTES 64
int gp_peer_alloc(int device, int64_t bytes, void** ptr_out, void* handle_out);
int gp_peer_free(int device, void* ptr);
int gp_peer_open(int device, const void* handle, void** ptr_out);
int gp_peer_close(int device, void* ptr);
int gp_peer_push(int device, int32_t n_dst, void* const* dst, const void* src_dev, int64_t bytes, void* stream);

/* gp_fbs_post with GpSolveExtras.wpu_dev added to sV (extras may be NULL; its other fields are ignored). */
int gp_fbs_post_ex(GpFeeder* f, int64_t S, int64_t ld, const double* spu_dev, const GpSolveExtras* extras,
                   const double* V_dev, const double* I_dev, double* sV_dev, double* Vmag_dev, double* Stx_dev,
                   double* Srx_dev, double* loss_dev, void* stream);

/* Number of rows of GpSolveExtras.gamma_dev for this feeder (0 = no regulators; negative = error). */
int gp_gamma_rows(GpFeeder* f);

/* gp_fbs_solve_post_batch with the optional inputs above (extras may be NULL).  Per-scenario ratios
 * and warm starts run on the streaming kernel. */
int gp_fbs_solve_batch_ex(GpFeeder* f, int64_t S, int64_t ld, const double* spu_dev, const GpSolveExtras* extras,
                          double* V_dev, double* I_dev, int32_t* iters_dev, int32_t* status_dev,
                          double* diff_dev, double* Vmag_dev, double* loss_dev, double tol, int32_t maxiter,
                          void* stream);

/*
 * Scenario studies (callers: a time-series driver looping over what the
 * reference does one `Circuit.set_kW / set_kvar` + `solve()` at a time, circuit.py:66-88).
 *
 * gp_der_rows: packed spu rows of an hourly profile x DER scenario study, built on the device.
 * Scenario first + s is hour h = idx % H of DER scenario d = idx / H (idx taken modulo H * D):
 *     spu[r][s] = m_h[h] * W0[r] - (Re W0[r] * mask[d][r]) * (pen[d] * pv_h[h])       (real part only)
 * with W0 [n_srows] complex the nominal load of each loaded bus-phase, m_h / pv_h [H] the hourly load
 * multiplier and PV shape, mask [D][n_srows] (0 / 1) the bus-phases that carry PV in DER scenario d and
 * pen [D] its penetration.  Each product and the difference are rounded separately, so the rows equal
 * the NumPy expression of bench.py:make_der_rows bit for bit.  All pointers are device pointers.
 *
 * gp_vmag_summary: per scenario min |V|, max |V| and the number of bus-phases with |V| outside
 * [vmin, vmax] (NaN counts as outside) over the rows of Vmag [n_vrows][ld]; any output may be NULL.
 */
int gp_der_rows(int device, int64_t S, int64_t ld, int64_t first, int32_t H, int32_t D, int32_t n_srows,
                const double* W0_dev, const double* m_h_dev, const double* pv_h_dev, const double* mask_dev,
                const double* pen_dev, double* spu_dev, void* stream);
/* gp_der_rows for the scenarios first + s * stride (s < S): the hour-k members of all chains of a chained
 * (warm-started, hour after hour) study, whose chains are runs of `stride` consecutive hours. */
int gp_der_rows_strided(int device, int64_t S, int64_t ld, int64_t first, int64_t stride, int32_t H, int32_t D,
                        int32_t n_srows, const double* W0_dev, const double* m_h_dev, const double* pv_h_dev,
                        const double* mask_dev, const double* pen_dev, double* spu_dev, void* stream);
int gp_vmag_summary(int device, int64_t S, int64_t ld, int32_t n_vrows, const double* Vmag_dev, double vmin,
                    double vmax, double* min_dev, double* max_dev, int32_t* count_dev, void* stream);


/* ------------------------------------------------------------------------
 * NR3 (three-phase Newton-Raphson), reference src/solution_nr3.py + src/nr3_lib/
 * ------------------------------------------------------------------------ */

/*
 * Flattened feeder for NR3.  Replaces the dense tensors SolutionNR3.__init__
 * builds (g_SB/b_SB :84-109, G_KVL :111-177, H/g/b :179-442): the same rows are
 * evaluated in closed form from these tables.  The unknown vector X is stored
 * PACKED as complex rows (A + jB per existing bus-phase, C + jD per existing
 * edge-phase): rows [0, n_vrows) are bus voltages, rows [n_vrows, n_vrows +
 * n_irows) edge currents.  Phases that do not exist are identity rows of J whose
 * solution is exactly 0 and are not stored.
 *
 * Step k is a non-slack bus with its incoming edge from its parent, parents before
 * children: a Line, a Transformer (Z = 0, :165-175) or the VoltageRegulators between one
 * pair of buses (GP_EDGE_VR: ratio rows and power-conservation rows, :444-515, with an
 * in-leg and an out-leg current per regulated phase).  A bus must carry exactly the
 * phases of its incoming edge.  The slack bus (index 0, solution.py:20) is the root.
 */
typedef struct {
    int32_t n_steps, n_vrows, n_irows, n_srows, n_child, n_lines;
    const int32_t* bus_vrow;    /* [n_steps][3] X row of (bus, ph) or -1               */
    const int32_t* par_vrow;    /* [n_steps][3] X row of (parent, ph) or -1            */
    const int32_t* edge_irow;   /* [n_steps][3] row (0-based in the I block) of the current ENTERING the bus */
    const int32_t* edge_orow;   /* [n_steps][3] row of the current LEAVING the parent: = edge_irow,
                                   except the out-leg of a regulator                   */
    const int32_t* load_srow;   /* [n_steps][3] spu row or -1                          */
    const int32_t* kind;        /* [n_steps] GP_EDGE_LINE (also transformers) or GP_EDGE_VR */
    const int32_t* child_begin; /* [n_steps]                                           */
    const int32_t* child_count; /* [n_steps]                                           */
    const int32_t* child_step;  /* [n_child] step index of each child bus              */
    const int32_t* child_irow;  /* [n_child][3] edge_orow of the child edges           */
    const double* z;            /* [n_steps][3][3][2] pu impedance (rmat/xmat, line.py:48-73) */
    const double* cappu;        /* [n_steps][3]                                        */
    const double* vr_gamma;     /* [n_steps][3] gamma per regulated phase, 0 elsewhere */
    int32_t root_vrow[3];
    int32_t root_load_srow[3];
    double root_cappu[3];
    double vslack[6];           /* solution.py:23-24, interleaved                      */
    double zip[3];              /* aPQ_p, aI_p, aZ_p (beta_S/I/Z = gamma_S/I/Z, :210-217) */
    double cA[3];               /* beta_Z + 0.5 beta_I hessian_mag[0][0] per phase (:252-258) */
    double cB[3];               /* beta_Z + 0.5 beta_I hessian_mag[1][1] per phase     */
    double f0_absent;           /* max |F| of the identity rows of absent phases at flat start */
    const int32_t* line_irow;   /* [n_lines][3] real lines, for Stx/Srx (map_output.py:40-55) */
    const int32_t* line_tx_vrow;
    const int32_t* line_rx_vrow;
} GpNr3Desc;

int gp_nr3_create(const GpNr3Desc* desc, int device, GpFeeder** out);

/* Bytes of device scratch gp_nr3_solve_batch needs for row pitch ld (negative = error). */
int64_t gp_nr3_workspace_bytes(GpFeeder* f, int64_t ld);

/*
 * Solve S scenarios from flat start; replaces SolutionNR3.solve()
 * (solution_nr3.py:592-632): X <- X - J^-1 F while max|F(previous X)| >= 1e-9
 * and it < maxiter, per scenario.
 *
 *   spu_dev    [n_srows][ld] complex   per-scenario bus loads (p + jq, load_group.py:25-37)
 *   X_dev      [n_vrows + n_irows][ld] complex   out: packed solution vector
 *   iters_dev  [S] int32               out: Newton iterations (the reference's local itercount)
 *   status_dev [S] int32               out: GP_ST_*
 *   fmax_dev   [S] double              out: max|F| seen by the last test
 */
int gp_nr3_solve_batch(GpFeeder* f, int64_t S, int64_t ld, const double* spu_dev, double* X_dev,
                       int32_t* iters_dev, int32_t* status_dev, double* fmax_dev,
                       void* workspace_dev, int64_t workspace_bytes, int32_t maxiter, void* stream);

/* gp_nr3_solve_batch with the optional per-scenario inputs of GpSolveExtras (extras may be NULL).
 * Per-scenario ratios and warm starts run on the thread-per-scenario kernel. */
int gp_nr3_solve_batch_ex(GpFeeder* f, int64_t S, int64_t ld, const double* spu_dev, const GpSolveExtras* extras,
                          double* X_dev, int32_t* iters_dev, int32_t* status_dev, double* fmax_dev,
                          void* workspace_dev, int64_t workspace_bytes, int32_t maxiter, void* stream);

/*
 * X -> results; replaces map_XNR / map_output (solution_nr3.py:634-663,
 * nr3_lib/map_output.py:4-79), including its 1e-12 scrub.  Any output may be NULL.
 *   V, sV, iNode [n_vrows][ld] complex; Vmag [n_vrows][ld] double;
 *   I, Stx, Srx  [n_irows][ld] complex (Stx/Srx: rows of real lines); loss [S] complex.
 */
int gp_nr3_map_output(GpFeeder* f, int64_t S, int64_t ld, const double* spu_dev, const double* X_dev,
                      double* V_dev, double* I_dev, double* Stx_dev, double* Srx_dev, double* sV_dev,
                      double* iNode_dev, double* Vmag_dev, double* loss_dev, void* stream);
/* gp_nr3_map_output with GpSolveExtras.wpu_dev added to sV (+ 1j * wpu, map_output.py:60) and through it to i_Node
 * (extras may be NULL; its other fields are ignored). */
int gp_nr3_map_output_ex(GpFeeder* f, int64_t S, int64_t ld, const double* spu_dev, const GpSolveExtras* extras,
                         const double* X_dev, double* V_dev, double* I_dev, double* Stx_dev, double* Srx_dev,
                         double* sV_dev, double* iNode_dev, double* Vmag_dev, double* loss_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GIGPOWER_B200_H */
