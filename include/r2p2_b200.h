/* r2p2_b200.h -- C ABI of the B200-native batched robot-step core for r2p2.
 *
 * This is the drop-in boundary for r2p2's one data-parallel hot path: a
 * whole population of robots x ticks is evaluated per call instead of one Python `Robot.update`
 * per robot per frame.  Plain C types only; caller owns host buffers, the library owns device
 * memory behind the opaque handle.  Every call returns 0 on success or a negative R2P2_E_* code;
 * the message is available from r2p2_last_error().  Nothing here ever falls back to the CPU: if
 * there is no CUDA device, r2p2_create fails.
 *
 * Reference interfaces replaced (paths under the r2p2 repository):
 *   r2p2_set_stage        <- utils.load_image + transpose          r2p2/utils.py:273-280, 302-303
 *   r2p2_set_robot        <- utils.create_robot / Robot.__init__   r2p2/utils.py:133-163, r2p2/robot.py:50-96
 *   r2p2_set_controller   <- controllers.load_controller + Neuro_controller.__build_network /
 *                            Inspyred_controller.set_network_params
 *                                                                  r2p2/controllers/controllers.py:20-39,
 *                                                                  r2p2/controllers/neurocontroller.py:106-127,
 *                                                                  r2p2/controllers/inspyred.py:132-144
 *   r2p2_upload_genomes   <- evolution.write_params -> ../res/weights.json -> set_network_params
 *                                                                  r2p2/evolution.py:24-27,
 *                                                                  r2p2/controllers/neurocontroller.py:139-160,168-178
 *   r2p2_reset            <- Controller.register_robot + episode reset
 *                                                                  r2p2/controllers/neurocontroller.py:77-85,100-104
 *   r2p2_run              <- the simulator loop `for r in robots: r.update(npdata, delta, True)`
 *                                                                  r2p2/utils.py:226-227, 233-234; r2p2/robot.py:389-402
 *   r2p2_read_fitness     <- Neuro_controller.fitness / __output_fitness -> ../res/fitness.txt -> evaluate_ann
 *                                                                  r2p2/controllers/neurocontroller.py:62-65,129-131,
 *                                                                  r2p2/evolution.py:29-52
 *   r2p2_read_state       <- Robot attributes x, y, orientation, speed, angular_velocity  r2p2/robot.py:63-75
 *   r2p2_read_trace       <- the per-tick sensor log (dst per ray)   r2p2/robot.py:337-352 and controller.dst
 */
#ifndef R2P2_B200_H
#define R2P2_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct r2p2_sim* r2p2_handle;

/* error codes */
#define R2P2_OK 0
#define R2P2_E_ARG (-1)      /* bad argument / call order      */
#define R2P2_E_CUDA (-2)     /* CUDA runtime error             */
#define R2P2_E_NODEV (-3)    /* no usable CUDA device          */
#define R2P2_E_LIMIT (-4)    /* exceeds a compiled-in limit    */

/* controller kinds (r2p2_set_controller) */
#define R2P2_CTRL_CONST 0    /* genome = [speed, angular_velocity]: naive_controller.py generalised   */
#define R2P2_CTRL_NEURO 1    /* neurocontroller.Neuro_controller: MLP, genome = flat C-order weights  */
#define R2P2_CTRL_LINEAR 2   /* inspyred.Inspyred_controller: [l_1..l_n, a_1..a_n, vmaxlin, vmaxang]  */
#define R2P2_CTRL_PID 3      /* pid_controller.Sequential_PID_Controller: genome = [ap, ai, ad, lp, li, ld]; goals: r2p2_set_pid */
#define R2P2_PID_MAX_GOALS 64 /* goal stack per robot: configured goals + temporary obstacle-avoid goals */

/* activations (sklearn names) */
#define R2P2_ACT_IDENTITY 0
#define R2P2_ACT_TANH 1
#define R2P2_ACT_RELU 2
#define R2P2_ACT_LOGISTIC 3

/* per-robot flag bits (r2p2_read_state) */
#define R2P2_F_COLLIDED 1u   /* Robot.collide() was called at least once                               */
#define R2P2_F_FAULT 2u      /* the reference would have raised (IndexError off-map, ValueError on NaN) */
#define R2P2_F_DONE 4u       /* evolve mode: the episode ended (timer ran out or collision)             */
#define R2P2_F_GOAL 16u      /* r2p2_set_goal: the goal predicate held after some tick                    */
#define R2P2_F_NOISE 32u     /* noise replay: the robot's stream ran out, later draws were 0.0 (results no longer follow the reference) */
#define R2P2_F_TRIGRANGE 8u  /* heading beyond the supported sin/cos range (|radians| >= 105414350)     */

/* limits */
#define R2P2_MAX_RAYS 64
#define R2P2_MAX_LAYERS 8
#define R2P2_MAX_WIDTH 64

/* lanes-per-robot choices for r2p2_set_mapping (0 = let the library choose from P) */
#define R2P2_LPR_AUTO 0

/* ---- lifetime ---------------------------------------------------------------------------- */
int r2p2_create(int device, r2p2_handle* out);
int r2p2_destroy(r2p2_handle h);
const char* r2p2_last_error(r2p2_handle h);            /* h may be NULL: error of the last failed create */
int r2p2_version(void);

/* Run everything on the caller's CUDA stream (a cudaStream_t passed as void*, e.g. torch's current
 * stream) instead of the handle's own non-blocking stream; NULL switches back. */
int r2p2_set_stream(r2p2_handle h, void* cuda_stream);
/* Measured FP64 issue peak of this device: a register-resident DFMA loop over all SMs; returns DFMA
 * thread-instructions per second (the denominator of the FP64-issue roofline). */
int r2p2_fp64_peak(r2p2_handle h, double* dfma_per_s);

/* ---- configuration ------------------------------------------------------------------------- */
/* Stage as the reference holds it after the transpose: levels[x*H + y], grey level 0..255
 * (0 = wall; 50 = the "crash" level tested at robot.py:249).  Packed on the device into
 * 1-bit planes (level==0, level==50). */
int r2p2_set_stage(r2p2_handle h, const uint8_t* levels_xy, int32_t W, int32_t H);
/* Same from the reference's own int32 array (npdata after flipud(rot90)). */
int r2p2_set_stage_i32(r2p2_handle h, const int32_t* levels_xy, int32_t W, int32_t H);

/* Robot json: sonar_range = [vr0, vr1], radius, max_speed, ray angles in degrees (list form, or the
 * int form already expanded by the host as range(0, 360, floor(365/n))). */
int r2p2_set_robot(r2p2_handle h, double vr0, double vr1, double radius, double max_speed,
                   const double* ray_deg, int32_t n_rays);

/* kind = R2P2_CTRL_*; NEURO: layers = [n_rays, hidden..., 2] (n_layers entries), activation = R2P2_ACT_*.
 * For CONST / LINEAR pass layers = NULL, n_layers = 0. */
int r2p2_set_controller(r2p2_handle h, int32_t kind, const int32_t* layers, int32_t n_layers, int32_t activation);

/* Sequential_PID_Controller on the device (r2p2/controllers/pid_controller.py:8-262): goal list, PID terms on heading and
 * distance, the obstacle-avoid state machine, backing off for ten ticks after a collision -- evaluated inside the step
 * kernel for every robot of the population; a robot's genome holds its six gains.  goals: [n_goals][2] (x, y), the same
 * list for every robot (per_robot = 0) or [P][n_goals][2] (per_robot = 1, P = the population size at the next reset).
 * max_acceleration: the controller's clamp (5 in the reference); acceleration: Robot.acceleration, which nothing in the
 * reference's tick changes (0).  The reference's register_robot moves the robot to goal[0]: pass that as the start pose
 * to r2p2_reset.  Goals are (re)loaded by r2p2_reset.  Needs at least three sonar rays. */
int r2p2_set_pid(r2p2_handle h, const double* goals, int32_t n_goals, int32_t per_robot, double max_acceleration, double acceleration);
/* Controller state per robot (any pointer may be NULL; arrays of P): goals left, state (0 advancing, 1 avoiding, 2 advancing
 * after avoiding), ticks of backing off left, target angle, and the current goal (NaN when the list is empty). */
int r2p2_read_pid(r2p2_handle h, int32_t* goals_left, int32_t* state, int32_t* backing, double* target_angle, double* goal_x, double* goal_y);

/* Sensor noise, Robot.has_noise / __generate_noise = np.random.normal(0, 0.5) (r2p2/robot.py:96,139-140,308-310): one draw
 * per unclamped sonar ray, in ray order.
 *   mode 0  off (the reference with has_noise = False)
 *   mode 1  replay: stream[P][len] holds each robot's draws, consumed sequentially -- with
 *           np.random.seed(s); np.random.normal(0, 0.5, len) a single robot reproduces the reference bit for bit
 *   mode 2  device generator: Philox4x32-10 keyed by `seed`, counter (draw index, robot), Box-Muller, sigma 0.5
 * Takes effect at the next r2p2_run; draw counters are zeroed by r2p2_reset. */
int r2p2_set_noise(r2p2_handle h, int32_t mode, uint64_t seed, const double* stream, int32_t P, int32_t len);
/* Draws consumed per robot since the last reset (uint32[P]). */
int r2p2_read_noise_draws(r2p2_handle h, uint32_t* ndraws);
/* Index of this handle's robot 0 in the whole (sharded) population: the device generator (mode 2) is keyed by the GLOBAL
 * robot index, so an individual draws the same noise whichever GPU evaluates it.  r2p2_ga_evaluate sets it to `first`. */
int r2p2_set_robot_index_offset(r2p2_handle h, uint32_t first);

/* Thread mapping: lanes per robot (1, 2, 4, 8, 16, 32) or R2P2_LPR_AUTO; stage_in_smem: 1 force the
 * shared-memory copy of the bit planes, 0 read them through L1/L2, -1 choose. */
int r2p2_set_mapping(r2p2_handle h, int32_t lanes_per_robot, int32_t stage_in_smem);

/* ---- population ---------------------------------------------------------------------------- */
/* genomes: [P][G] row-major doubles in HOST memory; copied to the device (async on the handle's stream,
 * the buffer may be reused when the call returns only if it is pageable; pinned buffers must stay
 * valid until the next synchronising call). */
int r2p2_upload_genomes(r2p2_handle h, const double* genomes, int32_t P, int32_t G);
/* Same, from a DEVICE pointer (e.g. a torch tensor's data_ptr on the same device). */
int r2p2_set_genomes_device(r2p2_handle h, const double* d_genomes, int32_t P, int32_t G);

/* Current population size and genome length (0, 0 before the first upload).  Every r2p2_read_* call copies P elements:
 * size the host arrays from this, not from a cached value (r2p2_ga_evaluate resizes the population to the shard). */
int r2p2_population_size(r2p2_handle h, int32_t* P, int32_t* G);

/* (Re)start every robot: pose from x0/y0/theta0 (n = 1: broadcast one pose, n = P: per robot),
 * speed = 0, odometry/origin = pose, distance = 0, flags = 0, controller timer = time_budget. */
int r2p2_reset(r2p2_handle h, const double* x0, const double* y0, const double* theta0, int32_t n, double time_budget);

/* Overwrite pose attributes of robots [first, first + count) the way host code assigns them in the reference
 * (Robot.set_position / set_last_position / `robot.orientation = ...`, r2p2/robot.py:98-107): only the arrays given
 * (any pointer may be NULL) are written, odometry, fitness and flags are left alone.  Host arrays of `count` doubles. */
int r2p2_write_pose(r2p2_handle h, int32_t first, int32_t count, const double* x, const double* y, const double* theta,
                    const double* last_x, const double* last_y);

/* ---- execution ----------------------------------------------------------------------------- */
/* Advance every robot by up to T ticks of Robot.update(env, delta) in ONE kernel launch (state stays in
 * registers across ticks).  delta_ctrl is what the controller subtracts from its timer per control()
 * call (utils.delta, in the controller's own unit).  evolve != 0: neurocontroller.py:72-85 episode
 * semantics -- a robot stops at the control() call where its timer is <= 0 (a collision zeroes the
 * timer) and its fitness is frozen there.  evolve == 0: free run for T ticks.  Asynchronous. */
int r2p2_run(r2p2_handle h, int32_t T, double delta, double delta_ctrl, int32_t evolve);
/* evolution.evaluate_ann for a whole population (r2p2/evolution.py:29-52: one candidate per call through weights.json /
 * fitness.txt) as ONE call: r2p2_upload_genomes + r2p2_reset + r2p2_run + r2p2_read_fitness, with the upload pipelined
 * when the population runs in three or more waves (a wave = as many robots as are resident at once): the first waves'
 * genomes go up, then the rest on a copy stream while the episode kernel already runs; the second piece has its own
 * launch on a second stream, so its CTAs move in as the first launch's retire.  Give
 * page-locked host memory for `genomes` to get the overlap (pageable memory works, the copies then serialise).
 * Arguments as for the four calls it replaces; fitness: [P], written before the call returns.  Results are identical
 * to the separate calls. */
int r2p2_evaluate(r2p2_handle h, const double* genomes, int32_t P, int32_t G, const double* x0, const double* y0, const double* theta0,
                  int32_t n, double time_budget, int32_t T, double delta, double delta_ctrl, int32_t evolve, double* fitness);
int r2p2_sync(r2p2_handle h);
/* Device time of the most recent r2p2_run kernel in milliseconds (CUDA events on the handle's stream). */
int r2p2_last_run_ms(r2p2_handle h, float* ms);

/* ---- results (each synchronises) ------------------------------------------------------------ */
int r2p2_read_fitness(r2p2_handle h, double* fitness /*[P]*/);
/* Any pointer may be NULL. Arrays of P elements. */
int r2p2_read_state(r2p2_handle h, double* x, double* y, double* theta, double* speed, double* omega,
                    uint32_t* flags, uint32_t* ncollide, uint32_t* ticks);
/* Totals over the population since the last reset: sonar-ray samples, collision-ray samples (pixels the
 * reference semantics test), executed robot-steps. */
int r2p2_read_counters(r2p2_handle h, uint64_t* sonar_samples, uint64_t* collision_samples, uint64_t* robot_steps);
/* Device pointer of the fitness vector [P] (for a collective issued by the caller, e.g. NCCL all-gather). */
int r2p2_fitness_device(r2p2_handle h, double** d_fitness);
/* Kernels launched by this handle since creation (every one is ours; no library kernels are used). */
int r2p2_launch_count(r2p2_handle h, uint64_t* n);

/* ---- checkpoint / resume ------------------------------------------------------------------------ */
/* Everything a running episode carries between launches (pose, last position, odometry, controller timer, fitness,
 * flags, collision / tick / noise-draw counters, sample counters, goal ticks) as one opaque host blob of
 * r2p2_state_bytes; loading it into a handle with the same population size, stage, robot and genomes continues the
 * episode bit for bit.  (The reference keeps all of this in Robot / Controller attributes, r2p2/robot.py:63-97,
 * r2p2/controllers/neurocontroller.py:27-33,100-104.) */
int r2p2_state_bytes(r2p2_handle h, uint64_t* bytes);
int r2p2_save_state(r2p2_handle h, void* blob);
/* The blob starts with a header (magic, format version, P, R, G, controller kind); a blob that does not fit this handle's
 * configuration is refused with R2P2_E_ARG. */
int r2p2_load_state(r2p2_handle h, const void* blob);

/* ---- per-tick trace (parity tests, Robot.update facade) -------------------------------------- */
/* Enable recording for the next runs: up to T_max ticks per run, layouts [T][P][...]. */
int r2p2_set_trace(r2p2_handle h, int32_t enable, int32_t T_max);
/* pose [T][P][5] = x, y, theta, speed, omega after the tick; dst [T][P][R] distances handed to the
 * controller; hitk [T][P][R] hit step or -1; hitpx [T][P][R][2] hit pixel or -1; nsamp [T][P][R] pixels
 * tested per ray; ncoll [T][P] collide() calls in the tick.  Any pointer may be NULL. */
int r2p2_read_trace(r2p2_handle h, int32_t T, double* pose, double* dst, int32_t* hitk, int32_t* hitpx,
                    uint32_t* nsamp, uint32_t* ncoll);
/* The collide() calls of each tick, for the collision lines of the robot log (Robot.collide, r2p2/robot.py:170-179,
 * called from __update_position, r2p2/robot.py:211-253), and the controller's answer, for the controller's own log
 * (log_step, r2p2/controllers/neurocontroller.py:97-98, inspyred.py:102-103): cinfo [T][P][6] = code, wx, wy, b, out0, out1.
 * code bit 0: wall collision at the proposed position (wx, wy) (robot.py:212); bits 1-4: map-border case 1..8 in
 * the order of robot.py:222-247 -- spot (W, b), (W, H), (W, 0), (0, b), (0, H), (0, 0), (b, H), (b, 0); bit 5: the
 * crash-radius / level-50 test fired and the robot went back to its last position, which is the spot (robot.py:249-252).
 * (out0, out1): what control() returned in that tick, before Robot.control's speed clamp (robot.py:385-387) -- for the
 * neuro controller (output 0 after its > 180 wrap, output 1); zeros for a tick in which control() did not get that far. */
int r2p2_read_trace_collisions(r2p2_handle h, int32_t T, double* cinfo);

/* ---- path planner's cost grid ------------------------------------------------------------------- */
/* path_planning.generate_grid (r2p2/path_planning.py:71-110) + Node.value (r2p2/node.py:64): the stage rasterised
 * into div_x x div_y cells.  values [div_y][div_x]: grey level at the cell centre, 0 if the cell's block contains a
 * wall pixel; costs: int(9 - value / 255 * 8), 1 (white) .. 9 (wall).  Either pointer may be NULL. */
int r2p2_cost_grid(r2p2_handle h, int32_t div_x, int32_t div_y, int32_t* values, int32_t* costs);

/* ---- goal flagging ---------------------------------------------------------------------------- */
/* Sequential_PID_Controller.is_at_goal (r2p2/controllers/pid_controller.py:207-213) as a per-run predicate for any
 * controller: after every completed tick, `env[int(gx), int(gy)] is 0 or norm(goal - pos) <= radius * 1.85`.  It
 * only sets R2P2_F_GOAL and records the tick; the dynamics never see it.  The goal pixel must be on the map
 * (Python index rules: -W <= int(gx) < W).  enable = 0 switches it off. */
int r2p2_set_goal(r2p2_handle h, int32_t enable, double gx, double gy);
/* goal_tick [P]: ticks completed when the predicate first held, 0xffffffff if it never did. */
int r2p2_read_goal(r2p2_handle h, uint32_t* goal_tick);

/* ---- controllers that run on the host (any reference Controller subclass: PID, PDDL, teleop ...) ---------- */
/* Robot.update is control() then the move (r2p2/robot.py:389-402); control() is check_sensors, the controller's
 * control(dst), the speed clamp (robot.py:373-387).  For a controller without a device form the tick is split at the
 * controller call:
 *   r2p2_sense  = Robot.check_sensors (robot.py:287-318) for every robot at its current pose: dst [P][R] distances
 *                 as the controller receives them (clamped, with noise if enabled), hitpx [P][R][2] hit pixel or -1
 *                 (may be NULL).  Poses are not touched.
 *   r2p2_move   = the rest of the tick with the controller's answer: commands [P][2] = (speed, angular_velocity) per
 *                 robot -> speed clamp, __update_angle, __update_position (robot.py:386-387, 257-267, 188-255).
 * Both are synchronous (host arrays).  The population is sized by r2p2_upload_genomes (any G) and r2p2_reset. */
int r2p2_sense(r2p2_handle h, double* dst, int32_t* hitpx);
int r2p2_move(r2p2_handle h, const double* commands, double delta);

/* ---- generation step of the EA on the device -------------------------------------------------- */
/* The reference leaves the EA to the user (r2p2/evolution.py:58-68 "TODO: Implement your EA here"; ga.py.gpg is
 * encrypted) and fixes the contract: generate_float candidates uniform in [min_gen, max_gen] (evolution.py:18-22),
 * evaluator(candidates) -> fitness, higher is better (evolution.py:29-52).  These calls keep population, fitness
 * and breeding (inspyred ec.GA operators: tournament selection, uniform crossover, Gaussian mutation clipped to
 * the range, elitism) in device memory.  All randomness is Philox4x32-10 of (seed, generation, individual, gene),
 * so every GPU of a multi-GPU run breeds the same next generation from the same fitness vector. */
/* Allocate a population of P individuals x G genes and draw generation 0 (generate_float). */
int r2p2_ga_create(r2p2_handle h, int32_t P, int32_t G, uint64_t seed, double min_gen, double max_gen);
/* tournament size 1..4, rates in [0, 1], elites 0..64.  Defaults: 2, 0.9, 0.05, 0.1, 1. */
int r2p2_ga_set_operators(r2p2_handle h, int32_t tournament, double crossover_rate, double mutation_rate,
                          double mutation_sigma, int32_t elites);
/* Device pointers of the current generation [P][G] and of its fitness vector [P] (target of the caller's
 * all-gather in a multi-GPU run). */
int r2p2_ga_population_device(r2p2_handle h, double** d_population);
int r2p2_ga_fitness_device(r2p2_handle h, double** d_fitness);
/* Host -> device fitness of individuals [first, first + count) (fitness computed elsewhere). */
int r2p2_ga_write_fitness(r2p2_handle h, const double* fitness, int32_t first, int32_t count);
/* Current generation and its fitness to the host; either pointer may be NULL. */
int r2p2_ga_read(r2p2_handle h, double* population, double* fitness);
/* evaluate_ann for individuals [first, first + count) of the current generation: they become this handle's
 * population, start from the pose / time budget of the last r2p2_reset(n = 1), run T ticks (r2p2_run), and their
 * fitness lands in the GA's fitness vector.  No host round trip. */
int r2p2_ga_evaluate(r2p2_handle h, int32_t first, int32_t count, int32_t T, double delta, double delta_ctrl, int32_t evolve);
/* Next generation from the current one and its fitness vector; the generation counter advances. */
int r2p2_ga_breed(r2p2_handle h);
/* Best individual of the generation last bred from: index, fitness, genome [G]; any pointer may be NULL
 * (the history_*.log record of evolution.py:45-47). */
int r2p2_ga_best(r2p2_handle h, int32_t* index, double* fitness, double* genome);
int r2p2_ga_generation(r2p2_handle h, uint32_t* generation);
/* Checkpoint of the EA: header (magic, version, P, G), seed, range, operators, generation counter, the current
 * generation [P][G] and its fitness vector [P].  r2p2_ga_load into a handle (with or without a GA of the same shape)
 * resumes the run reproducibly: the Philox draws depend only on (seed, generation, individual, gene). */
int r2p2_ga_state_bytes(r2p2_handle h, uint64_t* bytes);
int r2p2_ga_save(r2p2_handle h, void* blob);
int r2p2_ga_load(r2p2_handle h, const void* blob, uint64_t bytes);

/* ---- raw ray-caster (utils.search_edge_in_angle_with_limit, r2p2/utils.py:420-432) --------------- */
/* n independent rays against the current stage; host arrays in, host arrays out.
 * status: 1 hit, 0 miss, -1 fault (IndexError in the reference). */
int r2p2_cast_rays(r2p2_handle h, int32_t n, const double* ox, const double* oy, const double* angle_rad,
                   const double* limit, int32_t* status, double* hx, double* hy, int32_t* k, uint32_t* nsamp);

/* Same rays through the plain sample-by-sample restatement of the loop (no distance-field skipping): the
 * in-library cross-check of r2p2_cast_rays. */
int r2p2_cast_rays_plain(r2p2_handle h, int32_t n, const double* ox, const double* oy, const double* angle_rad,
                         const double* limit, int32_t* status, double* hx, double* hy, int32_t* k, uint32_t* nsamp);

#ifdef __cplusplus
}
#endif
#endif /* R2P2_B200_H */
