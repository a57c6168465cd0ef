/* sapio.h -- C-ABI of the B200-native Sapiogenesis world step.
 *
 * The reference (edgecase963/Sapiogenesis) has no FFI: its per-tick path is the Python call chain
 *   Environment.update            physics.py:425-442
 *     -> sprites.update_organisms sprites.py:1809-1856   (rules: Organism.update :1627-1681, _update_cell :1470-1606,
 *                                                         neural.activate neural.py:91-366, neural.train_network :369-423,
 *                                                         reproduction sprites.py:1441-1459)
 *     -> pymunk.Space.step(0.06)  physics.py:432         (Chipmunk2D: broadphase, narrowphase, sequential impulses)
 *     -> Sprite.updateSprite      physics.py:250-275     (viscous friction)
 *     -> updateUI                 Sapiogenesis.py:57-65  (gas clamps, population recount)
 * Each entry point below replaces one stage of that chain and cites it. All pointers are raw device
 * pointers (CUDA library) or host pointers (oracle); the caller (PyTorch) owns every buffer including
 * scratch. No torch types, no exceptions, no allocation, no ownership transfer across this boundary.
 * Return value: 0 = ok, negative = SAPIO_E_*; message via sapio_last_error().
 */
#ifndef SAPIO_H
#define SAPIO_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SAPIO_ABI_VERSION 1

/* ---- fixed tile sizes ---- */
#define SAPIO_MAXC   64                         /* cell rows per organism slot */
#define SAPIO_NPAIR  (SAPIO_MAXC * (SAPIO_MAXC - 1) / 2)
#define SAPIO_ICAP   192                        /* intra-organism contact slots per organism */
#define SAPIO_MAXL   12                         /* linear layers per brain (input + <=10 hidden + output) */
#define SAPIO_NWALL  4
#define SAPIO_HIST   10                         /* neural.stimulation_memory (neural.py:34) */
#define SAPIO_STM    30                         /* short-term memory cap (neural.py:315) */
#define SAPIO_MEMROWS 232                       /* memory_limit 200 (+30 appended before the trim, neural.py:392-404) */
#define SAPIO_RNN_MAXH 16                       /* largest hidden_rnn_size (physics.py:308 default 6) */
#define SAPIO_HM 90                             /* dna.hidden_memory: <= 30 own entries + <= 60 inherited ones (sprites.py:386 is not erased at birth) */

/* ---- cell types: sprites.cell_types (sprites.py:22) + "heart" (sprites.py:573) ---- */
enum {
    SAPIO_T_BARRIER = 0, SAPIO_T_CARNIV = 1, SAPIO_T_EYE = 2, SAPIO_T_OLFACTORY = 3,
    SAPIO_T_CO2C = 4, SAPIO_T_PUSH = 5, SAPIO_T_BODY = 6, SAPIO_T_ROTATE = 7, SAPIO_T_HEART = 8,
    SAPIO_NTYPES = 9
};

/* c_alive values */
enum { SAPIO_CELL_ABSENT = 0, SAPIO_CELL_ALIVE = 1, SAPIO_CELL_DYING = 2 /* died this tick, body not yet moved to the dead pool */ };

/* org_flags bits */
enum { SAPIO_F_IMMORTAL = 1, SAPIO_F_MALADAPTIVE = 2, SAPIO_F_FINITE_MEMORY = 4, SAPIO_F_MIRROR_X = 8, SAPIO_F_MIRROR_Y = 16 };

/* org_flags bits (bit0 immortal, bit1 maladaptive, bit2 finite_memory, bit3 mirror_x, bit4 mirror_y: sapio_fields.def) */
#define SAPIO_F_GHOST 32          /* slab mode: a copy of an organism another slab owns */

/* org_req bits */
enum { SAPIO_REQ_REPRODUCE = 1, SAPIO_REQ_THOUGHT = 2, SAPIO_REQ_TRAIN = 4 };

/* g_stats slots */
enum {
    SAPIO_STAT_BIRTHS = 0, SAPIO_STAT_DEATHS = 1, SAPIO_STAT_CELL_DEATHS = 2, SAPIO_STAT_DEAD_REMOVED = 3,
    SAPIO_STAT_THINKS = 4, SAPIO_STAT_TRAINS = 5, SAPIO_STAT_XCONTACTS = 6, SAPIO_STAT_ICONTACTS = 7,
    SAPIO_STAT_BIRTH_FAIL_PLACE = 8, SAPIO_STAT_BIRTH_FAIL_CAPS = 9, SAPIO_STAT_OVERFLOW = 10,
    SAPIO_STAT_ORG_TICKS = 11, SAPIO_STAT_LEVELS = 12, SAPIO_STAT_COUPLED = 13,
    SAPIO_STAT_RARE_PATHS = 14,    /* bits: SAPIO_RARE_* -- branches only large or aged worlds reach (set, never cleared: test evidence) */
    SAPIO_STAT_FULL_SCANS = 15     /* birth placements that scanned every organism (more than REPRO_NEAR_CAP neighbours within reach) */
};

enum { SAPIO_RARE_GRID_LEVELS = 1 /* cooperative kernel relaxed contact levels grid-wide */, SAPIO_RARE_PACKED_COUPLED = 2 /* a component too large for any CTA tier went to the cooperative kernel */, SAPIO_RARE_FULL_SCAN = 4, SAPIO_RARE_TRAIN_QUEUE = 8, SAPIO_RARE_BIRTH_CAP = 16 };

/* phase mask for sapio_tick */
enum {
    SAPIO_PH_RULES = 1, SAPIO_PH_THINK = 2, SAPIO_PH_TRAIN = 4, SAPIO_PH_REPRODUCE = 8,
    SAPIO_PH_STEP = 16, SAPIO_PH_FRICTION = 32, SAPIO_PH_POST = 64, SAPIO_PH_ALL = 127,
    /* caller's assertion: no body was added or moved since the last SAPIO_PH_STEP on this workspace, so the spatial grid that
     * step built (positions after integration) is still the grid of the world and neural.activate's sensors may use it */
    SAPIO_PH_REUSE_GRID = 128
};

enum { SAPIO_OK = 0, SAPIO_E_ARG = -1, SAPIO_E_CUDA = -2, SAPIO_E_WORKSPACE = -3, SAPIO_E_OVERFLOW = -4, SAPIO_E_NODEVICE = -5 };

/* Frozen world parameters. Defaults and their reference sources: SURVEY.md Appendix B. */
typedef struct SapioParams {
    /* capacities */
    int32_t n_org, n_dead, x_cap, in_cap, brain_cap, adam_cap, width_cap;
    int32_t rnn_cap;          /* capacity of the recurrent-state arrays (RNNH of sapio_fields.def): 0 = a world without RNN brains, else <= SAPIO_RNN_MAXH */
    /* world: environment_settings.json:1, physics.py:280-345 */
    float width, height, frame_width, wall_elast, wall_fric;
    float dt_wall;            /* update_speed/1000 = 0.030 s: the uDiff of every rule (virtual clock) */
    float dt_phys;            /* Environment.worldSpeed = 0.06 (physics.py:319-320) */
    float friction_pct;       /* Environment.friction = 60 %/s (physics.py:283) */
    /* cadences in ticks (virtual clock; sprites.py:41,53,55,62) */
    int32_t think_ticks, train_ticks, repro_ticks, age_ticks;
    /* sprites.py:37-65 */
    float co2_ratio, dispersion_rate, heal_rate, mass_cutoff, break_damage, starvation_rate;
    float max_push, max_rot, eyesight_mult, olfactory_mult;
    /* Environment.info (physics.py:298-309), sprites.py:43 */
    float mutation_severity, brain_mutation_severity, learning_rate;
    int32_t population_limit, offspring_amount, weight_persistence, memory_limit;
    /* Chipmunk space defaults */
    int32_t iterations;
    float collision_slop, collision_bias, joint_error_bias;
    /* spawn recipe: add_creature_clicked widget defaults (Sapiogenesis.py:297-320); the size and
     * mass ranges are also what mutate_cell_count re-uses (sprites.py:937-940) */
    int32_t cell_lo, cell_hi, size_lo, size_hi, mass_lo, mass_hi, hidden_lo, hidden_hi;
    float mirror_px, mirror_py;
    /* RNG */
    uint64_t seed;
    /* grid */
    float grid_cell;          /* broadphase cell edge (>= 2 * max solid radius) */
    int32_t grid_nx, grid_ny;
    /* spatial slabs (second mode): this world is the slab [slab_x0, slab_x1) of a larger one, in the larger
     * world's coordinates. slab_mode = 0: an ordinary whole world. In slab mode organisms flagged SAPIO_F_GHOST (copies of the
     * neighbours' boundary organisms) take no part in the gas exchange, do not reproduce and are not counted; the global gas
     * scalars and the population are composed across the slabs by the host (sapiogenesis_b200/slabs.py). */
    int32_t slab_mode;
    float grid_x0;            /* x of the first grid column (0 for a whole world) */
    float slab_x0, slab_x1;
    /* Environment.info["use_rnn"] / ["hidden_rnn_size"] (physics.py:305,308) as ONE number: the hidden-state size of the brains built from
     * now on (spawn, birth; sprites.py:1683-1690), 0 = feed-forward networks.Network. 0 or rnn_cap: the reference cannot change the size
     * while RNN organisms live either (torch.tensor(trainingHidden) of ragged rows raises, neural.py:412). */
    int32_t rnn_hidden;
} SapioParams;

/* The world: one raw pointer per SoA array of include/sapio_fields.def. */
typedef struct SapioWorld {
    SapioParams p;
#define SAPIO_FIELD(name, ctype, scope, per) ctype *name;
#include "sapio_fields.def"
#undef SAPIO_FIELD
} SapioWorld;

/* body id space used for contact keys and canonical ordering */
static inline uint32_t sapio_body_cell(const SapioParams *p, int org, int k) { (void)p; return (uint32_t)(org * SAPIO_MAXC + k); }
static inline uint32_t sapio_body_dead(const SapioParams *p, int d) { return (uint32_t)(p->n_org * SAPIO_MAXC + d); }
static inline uint32_t sapio_body_wall(const SapioParams *p, int w) { return (uint32_t)(p->n_org * SAPIO_MAXC + p->n_dead + w); }
/* index of joint (i,j), i<j, in lexicographic order == Chipmunk constraint insertion order (sprites.py:1721-1726) */
static inline int sapio_pair_index(int i, int j) { return i * (2 * SAPIO_MAXC - i - 1) / 2 + (j - i - 1); }

/* ------------------------------------------------------------------------------------------------
 * CUDA library entry points (libsapio_b200.so). `stream` is a cudaStream_t passed as void*.
 * ---------------------------------------------------------------------------------------------- */
int sapio_abi_version(void);
const char *sapio_last_error(void);
int sapio_device_count(void);

/* scratch the caller must provide to every call below (grid tables, sort buffers, adjacency, levels) */
size_t sapio_workspace_bytes(const SapioParams *p);

/* One whole world tick == Environment.update (physics.py:425-442) with the hooks bound as at
 * Sapiogenesis.py:687-688. `phases` = SAPIO_PH_* mask (SAPIO_PH_ALL for the full step). */
int sapio_tick(const SapioWorld *w, void *ws, size_t ws_bytes, uint32_t phases, int n_ticks, void *stream);

/* Stages, callable on their own (parity tests drive these one by one):                               */
/* pymunk.Space.step(dt_phys)                                                          physics.py:432  */
int sapio_space_step(const SapioWorld *w, void *ws, size_t ws_bytes, void *stream);
/* Sprite.updateSprite for every solid sprite                                          physics.py:250  */
int sapio_friction(const SapioWorld *w, void *stream);
/* updateUI's simulation-relevant part: gas clamps, population recount                 Sapiogenesis.py:57-65 */
int sapio_post(const SapioWorld *w, void *ws, size_t ws_bytes, void *stream);
/* sprites.update_organisms: per-organism rules incl. neural.activate when SAPIO_PH_THINK is set          sprites.py:1809, neural.py:91 */
int sapio_rules(const SapioWorld *w, void *ws, size_t ws_bytes, uint32_t phases, void *stream);
/* neural.train_network for every organism sapio_rules marked as due (org_req & SAPIO_REQ_TRAIN)           neural.py:369 */
int sapio_train(const SapioWorld *w, void *ws, size_t ws_bytes, void *stream);
/* deferred carnivore damage / eating, dying cells -> dead bodies, gradual_dispersion, bounds              sprites.py:1326,1497,1752,1849 */
int sapio_settle(const SapioWorld *w, void *ws, size_t ws_bytes, void *stream);
/* Organism.reproduce / DNA.mutate / get_new_organism_position / build_body for org_req & SAPIO_REQ_REPRODUCE  sprites.py:1441,1068,194,1698 */
int sapio_reproduce(const SapioWorld *w, void *ws, size_t ws_bytes, void *stream);
/* add_creature_clicked: DNA().randomize(spawn ranges) + Organism + build_body into the given free slots at the given
 * positions (device arrays: n slots, 2n doubles); uids are g_next_uid, g_next_uid + 1, ...               Sapiogenesis.py:297-323 */
int sapio_spawn(const SapioWorld *w, void *ws, size_t ws_bytes, const int32_t *d_slots, const double *d_pos, int n, void *stream);
/* neural.activate(network, environment, organism, uDiff) for n given organisms now (device array of slots): sensors,
 * forward pass, boredom gate, history / short-term memory, movement outputs, stimulation -- outside the think cadence   neural.py:91-366 */
int sapio_activate(const SapioWorld *w, void *ws, size_t ws_bytes, const int32_t *d_orgs, int n, void *stream);
/* neural.train_network(organism, epochs=1) for n given organisms now: memory curation + one Adam epoch                neural.py:369-423 */
int sapio_train_network(const SapioWorld *w, void *ws, size_t ws_bytes, const int32_t *d_orgs, int n, void *stream);
/* parity probe: sprite.info["view"] of one eye / olfactory cell as ascending body ids (device out, count may exceed cap) sprites.py:221-236 */
int sapio_sensor_view(const SapioWorld *w, void *ws, size_t ws_bytes, int org, int cell, uint32_t *d_out, int cap, int32_t *d_count, void *stream);

/* ---- island migration (one world per GPU). The reference has one world; this is the exchange step of the
 * multi-GPU extension: a record is every per-organism array (scopes ORG / CELL / PAIR / ICON of sapio_fields.def) of one
 * organism, i.e. its DNA tables (sprites.py:329-408), cells, brain (networks.py:137-230), optimizer state and memories. */
size_t sapio_organism_record_bytes(const SapioParams *p);
/* gathers the organisms in d_slots into n consecutive records */
int sapio_pack_organisms(const SapioWorld *w, const int32_t *d_slots, int n, void *d_buf, size_t buf_bytes, void *stream);
/* scatters n records into the (free) slots d_slots; the arrivals get the uids g_next_uid, g_next_uid + 1, ... */
int sapio_unpack_organisms(const SapioWorld *w, const int32_t *d_slots, int n, const void *d_buf, size_t buf_bytes, void *stream);
/* frees the slots of organisms that left (no dead cells remain: they live on in another world) */
int sapio_release_organisms(const SapioWorld *w, const int32_t *d_slots, int n, void *stream);

/* Island migration without a host round trip (wire format: include/sapio_migrate.h; order of every choice: csrc/migrate2.cuh).
 * _out: picks up to `count` living organisms (seeded walk), packs their ragged records into d_buf until it is full, frees their
 *       slots and scrubs what pointed at them. _in: places every record of d_buf by the reference's rule for a new organism
 *       (box inside the world, blocks no living organism's box: viable_organism_position, sprites.py:156-192; first viable of 360
 *       seeded candidates), scatters it into a free slot with a fresh uid, removes dead cells inside its box (sprites.py:184-190).
 * Arrivals that find no place or no free slot are dropped and counted. The buffer has a fixed size (_buffer_bytes), so the
 * transport (NCCL send/recv) needs no size exchange.
 * _counts (synchronises the stream): {chosen, packed, accepted, dropped} of the last calls, then totals {out, in, dropped, stayed}. */
/* Spatial slabs of one large world (csrc/halo.cuh; driver: sapiogenesis_b200/slabs.py). The world step of a slab is the ordinary
 * one; what the slabs exchange is (a) every tick, the organisms (ragged records) and dead cells next to a boundary -- _halo_pack /
 * _halo_unpack / _halo_sweep, transport by the host (NCCL) -- and (b) three scalars composed in rank order: the slabs' gas maps
 * (the reference updates co2 / oxygen cell by cell, sprites.py:1409-1439: the maps do not commute), refunds, dispersion products.
 * sapio_rules_part: part 1 = begin + gas scan + think, 2 = main (+ finish outside slab mode), 3 = both (== sapio_rules). */
int sapio_rules_part(const SapioWorld *w, void *ws, size_t ws_bytes, uint32_t phases, int part, void *stream);
int sapio_post_advance(const SapioWorld *w, void *ws, size_t ws_bytes, void *stream);
int sapio_slab_offsets(const SapioWorld *w, void *ws, size_t ws_bytes, int64_t *out4);   /* byte offsets in ws: gas map (2 doubles), refund sum, dispersion product, owned population (int) */
int sapio_halo_pack(const SapioWorld *w, void *ws, size_t ws_bytes, float x_lo, float x_hi, void *d_buf, size_t buf_bytes, void *d_dead, size_t dead_bytes, void *stream);
int sapio_halo_unpack(const SapioWorld *w, void *ws, size_t ws_bytes, int side, void *d_buf, size_t buf_bytes, void *d_dead, size_t dead_bytes, void *stream);
int sapio_halo_sweep(const SapioWorld *w, void *ws, size_t ws_bytes, void *stream);
int sapio_slab_gas_begin(const SapioWorld *w, void *ws, size_t ws_bytes, const double *d_maps, int n, int rank, void *stream);
int sapio_slab_gas_finish(const SapioWorld *w, const double *d_maps, int n, const double *d_refunds, void *stream);
int sapio_slab_disperse_finish(const SapioWorld *w, const double *d_prods, int n, void *stream);
int sapio_slab_population(const SapioWorld *w, const int32_t *d_counts, int n, void *stream);

size_t sapio_migrate_buffer_bytes(const SapioParams *p, int count, int bytes_per_organism);
int sapio_migrate_out(const SapioWorld *w, void *ws, size_t ws_bytes, int count, int rank, void *d_buf, size_t buf_bytes, void *stream);
int sapio_migrate_in(const SapioWorld *w, void *ws, size_t ws_bytes, void *d_buf, size_t buf_bytes, void *stream);
int sapio_migrate_counts(const SapioWorld *w, void *ws, size_t ws_bytes, int32_t *h_out8, void *stream);

/* ---- end-to-end path for a GUI host (host buffers; the copies are inside the call) --------------------------------
 * Uploads the per-tick event block updateUI applies (Sapiogenesis.py:37-47), runs n_ticks ticks, then downloads what the
 * reference pushes to Qt every frame: the UI counters (Sapiogenesis.py:62-69) and the pose of every solid sprite
 * (Sprite.setPos / setRotation, physics.py:253-272), packed. Host pointers should be page-locked. Synchronises `stream`. */
typedef struct SapioControl { float drought, poison, random_cells; int32_t pad; } SapioControl;   /* amounts / probability per tick = severity * uDiff */
typedef struct SapioFrameHeader {
    double co2, o2; int64_t tick; int32_t population, n_cells, n_dead, pad; int64_t stats[16];
} SapioFrameHeader;
int sapio_tick_host(const SapioWorld *w, void *ws, size_t ws_bytes, uint32_t phases, int n_ticks,
                    const SapioControl *h_control,      /* host, in  */
                    SapioFrameHeader *h_header,         /* host, out */
                    float *h_cells, int cell_cap,       /* host, out: n_cells x (x, y, angle, organism slot) */
                    float *h_dead, int dead_cap,        /* host, out: n_dead  x (x, y, angle, mass)          */
                    void *stream);
/* The same exchange, pipelined: _begin queues the ticks, the packing and -- on a copy stream of the library -- the download, and
 * returns at once; _end waits for that download (and fetches a remainder if the world grew past the size guessed from the previous
 * frame). A host that calls begin(k + 1) before end(k) with two sets of pinned buffers overlaps the download of frame k with the
 * computation of tick k + 1. One frame in flight per device. */
int sapio_tick_host_begin(const SapioWorld *w, void *ws, size_t ws_bytes, uint32_t phases, int n_ticks, const SapioControl *h_control,
                          SapioFrameHeader *h_header, float *h_cells, int cell_cap, float *h_dead, int dead_cap, void *stream);
int sapio_tick_host_end(void);

/* the same events as a stage (applied after the step, before sapio_post): drought and poison amounts, and the probability that
 * one piece of food (add_dead_clicked: size 30, mass 12) appears at a random place this tick   Sapiogenesis.py:37-53, 490-507 */
int sapio_events(const SapioWorld *w, void *ws, size_t ws_bytes, float drought, float poison, float random_cells, void *stream);

/* ---- measurement hooks ------------------------------------------------------------------------------------------- */
/* per-stage device time (CUDA events on the launching stream); names: sapio_profile_name(i), i < sapio_profile_stages() */
int sapio_profile(int enable);
int sapio_profile_stages(void);
const char *sapio_profile_name(int i);
int sapio_profile_read(double *ms_out, int64_t *count_out, int reset);
/* number of kernels of this library launched so far in this process */
uint64_t sapio_launch_count(void);
/* Organisms in contact with other bodies ("coupled") are solved by one of two cooperative kernels with identical results: one
 * warp per organism, or lane-packed size classes. The packed one is launched when the previous tick counted more than `n`
 * coupled organisms (default 2400; -1 = always, INT_MAX = never). Returns the previous threshold. Tuning / test knob. */
int sapio_set_coupled_pack_from(int n);
/* Diagnostic builds (-DSAPIO_DEBUG_FLAGS) honour A/B switches (bit 0: solver classes by rows in use, bit 1: library sqrt / divide in the
 * pin prestep); the production build ignores them. Returns the previous value. */
int sapio_debug_flags(int flags);
/* Residency of the organism solver, per size class cls (cell rows <= 8 (cls + 1)): out6 = {lanes per organism, organisms per warp,
 * warps per CTA, shared-memory bytes per warp, resident CTAs on the device (0 before the first step), SMs}. Measurement hook: the
 * solver is bound by a dependent chain of wavefront steps per organism times the organisms in flight, not by memory. */
int sapio_solver_info(int cls, int32_t *out6);

#ifdef __cplusplus
}
#endif
#endif /* SAPIO_H */
