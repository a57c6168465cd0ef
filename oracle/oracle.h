/* oracle.h -- internal declarations of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY. This directory is a CPU restatement of the reference's per-tick world step
 * (edgecase963/Sapiogenesis: physics.py, sprites.py, neural.py, networks.py, and the Chipmunk2D 7.0.x
 * subset that pymunk>=6.2.1 executes for it -- pymunk is an un-vendored dependency, requirements.txt:2,
 * absent from the reference checkout and from this image). Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it. The product (sapiogenesis_b200/) never does.
 *
 * Pinning status (see DESIGN.md "Oracle"):
 *   - brain forward / train epoch / eye+olfactory encoding / stimulation / memory curation / finish_cell_info /
 *     DNA.randomize / DNA.mutate / get_rect / placement predicate: PINNED against the reference's own Python
 *     (networks.py, neural.py, sprites.DNA imported in the build container; vectors in tests/golden/).
 *   - Space.step: PARITY UNPINNED at the Chipmunk boundary (no pymunk, no Chipmunk source, no reference tests);
 *     restated from Chipmunk2D 7.0.3's published algorithm (cpSpaceStep.c, cpArbiter.c, cpPinJoint.c,
 *     cpCollision.c) and checked against closed-form cases only.
 *
 * Arithmetic: state arrays are fp32 (identical bytes to the GPU's); the oracle computes in double like the
 * reference (Python float / cpFloat), except the brain (fp32, as torch) and the places marked MIRROR32 where the
 * single fp32 operation the GPU performs is reproduced so that geometric predicates see identical positions.
 */
#ifndef SAPIO_ORACLE_H
#define SAPIO_ORACLE_H

#include "../include/sapio.h"
#include "philox.h"

#define MAXC SAPIO_MAXC

/* uniform source: either a Philox stream or a caller-provided array (pinning tests feed the reference the same array) */
typedef struct USrc {
    const float *arr; int n_arr; int pos;   /* array mode when arr != NULL */
    SapioStream st;                          /* stream mode otherwise */
    int overrun;
} USrc;
double usrc_next(USrc *u);

/* genome under construction == sprites.DNA (sprites.py:329-408), dict order made explicit */
typedef struct OGenome {
    int n;                       /* cells, in dna.cells insertion order */
    int id[MAXC];                /* dna cell ids (1-based) */
    int type[MAXC];
    int first[MAXC];
    int mirror_id[MAXC];         /* 0 = None */
    int parent_id[MAXC];         /* 0 = grows from nothing (first cell) */
    double size[MAXC], mass[MAXC], elast[MAXC], fric[MAXC];
    double relx[MAXC], rely[MAXC], growx[MAXC], growy[MAXC];
    double maxhealth[MAXC], use_idle[MAXC], use_act[MAXC], storage[MAXC], damage[MAXC];
    int mirror_x, mirror_y;
    double curiosity;
    int n_hidden;                /* len(brain_structure["hidden_layers"]) */
    int hidden[16];
    float *hw;                   /* hidden weights, layer l: [hidden[l+1]][hidden[l]] row-major, concatenated */
    int hw_len;
    int generation;
} OGenome;

void og_init(OGenome *g);
void og_free(OGenome *g);
void og_copy(OGenome *dst, const OGenome *src);
int  og_index_of(const OGenome *g, int id);
void og_finish_cell_info(OGenome *g, int j);
int  og_randomize(OGenome *g, const SapioParams *p, USrc *u, USrc *u_init);
int  og_mutate(OGenome *child, const OGenome *parent, const SapioParams *p, double severity, double neural_severity,
               int weight_persistence, USrc *u, USrc *u_fill, USrc *u_garb);
void og_get_rect(const OGenome *g, double r[4]);
void og_get_size(const OGenome *g, double s[2]);
int  og_instance_order(const OGenome *g, int order[MAXC]);
int  og_input_size(const OGenome *g);
int  og_brain_params(const OGenome *g);
int  og_fits_caps(const OGenome *g, const SapioParams *p);

/* world <-> genome */
void ow_genome_from_world(const SapioWorld *w, int org, OGenome *g);
int  ow_materialize(SapioWorld *w, int org, const OGenome *g, double px, double py, int64_t uid, int32_t tick, USrc *u_init);
void ow_inherit_recurrent(SapioWorld *w, int parent, int child);
void ow_org_rect(const SapioWorld *w, int org, double r[4]);
int  ow_viable_position(const SapioWorld *w, const OGenome *g, double px, double py);

/* stages (oracle side of include/sapio.h's stage list) */
int oracle_space_step(SapioWorld *w);
int oracle_friction(SapioWorld *w);
int oracle_rules(SapioWorld *w, uint32_t phases);
int oracle_train(SapioWorld *w);
int oracle_settle(SapioWorld *w);
int oracle_reproduce(SapioWorld *w);
int oracle_post(SapioWorld *w);
int oracle_tick(SapioWorld *w, uint32_t phases, int n_ticks);

/* small shared helpers */
static inline double og_sq(double x) { return x * x; }

#endif
