/* qmo_priv.h -- what the oracle's translation units share behind qmo.h (TEST INFRASTRUCTURE) */
#ifndef QMO_PRIV_H
#define QMO_PRIV_H
#include "qmo.h"
struct qmo_ref {
    int n_contigs, k;
    int64_t l_pac, *off, *len;
    uint8_t *fwd;
    int64_t n_km;
    uint64_t *km_key;
    uint32_t *km_pos;
    const void *fm;               /* qmo_fm_t of the same genome (not owned), or NULL: see qmo_ref_set_fm */
    int fm_max_mem_intv;
};
/* qmo_fm.c */
int qmo_fm_seeds(const void *F, const qmo_ref_t *R, const qmo_opt_t *o, int len, const uint8_t *q, int max_mem_intv,
                 int64_t *seeds /* 3 per seed: rbeg, qbeg, len */, int max_seeds);
#endif
