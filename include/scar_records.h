/*
 * scar_records.h — per-read result record shared by the CUDA path, the C-ABI
 * and the CPU oracle (tests compare the two byte for byte).
 *
 * One record is what the reference's per-read call returns, in fixed width:
 *   SlidingWindow.sliding_window (reference scarmapper/SlidingWindow.pyx:20-171)
 *   returns ([], summary) or ([ldel, rdel, ins, mh, read, clj, crj, tlj, trj, hr], summary);
 *   every string in that list is a slice of the read or of the target, so the
 *   four junction coordinates + status + flags reproduce it exactly:
 *     ldel = target[tlj:cut]      rdel = target[cut:trj]
 *     ins  = read[clj:crj]  (flag INSERTION)     mh = read[crj:clj] (flag MICROHOMOLOGY)
 *   and the summary_data increments (INDEL_Processing.py:94,147-169;
 *   SlidingWindow.pyx:59,97,109-112,145-167) are a function of status/flags/hr_n.
 */
#ifndef SCAR_RECORDS_H
#define SCAR_RECORDS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status */
enum {
    SCAR_ST_UNASSIGNED  = 0, /* no sample index matched (INDEL_Processing.py:1194-1198) */
    SCAR_ST_FILTERED    = 1, /* N fraction / min length filter (INDEL_Processing.py:147-154) */
    SCAR_ST_NO_JUNCTION = 2, /* SlidingWindow.pyx:96-98   summary[6][0]++ */
    SCAR_ST_NO_CUT      = 3, /* SlidingWindow.pyx:58-60,166-168  summary[6][1]++ */
    SCAR_ST_N_INSERTION = 4, /* SlidingWindow.pyx:140-141 dropped, no counter */
    SCAR_ST_SCAR        = 5  /* 10-field list returned */
};

/* flags */
enum {
    SCAR_FL_LEFT_DEL  = 1,  /* tlj < cutsite   summary[2]++ (SlidingWindow.pyx:148-150) */
    SCAR_FL_RIGHT_DEL = 2,  /* trj > cutsite   summary[3]++ (:153-155) */
    SCAR_FL_INSERTION = 4,  /* 0 < clj < crj   summary[4]++ (:136-145) */
    SCAR_FL_MICROHOM  = 8,  /* clj > crj > 0   summary[5]++ (:159-163) */
    SCAR_FL_HR        = 16, /* donor seen, label "HR" (:108-117) */
    SCAR_FL_CUTWINDOW = 32  /* NO_CUT taken through the cutwindow test (:58-60) */
};

#define SCAR_SAMPLE_NONE 0xFFFFu
#define SCAR_PHASE_NONE  (-1) /* "No Read N Phasing" (INDEL_Processing.py:972-975) */
#define SCAR_PHASE_NA    (-2) /* phasing not evaluated (not Illumina / unassigned / empty table) */

typedef struct scar_record {
    uint8_t  status;
    uint8_t  flags;
    uint16_t sample;   /* manifest position of the sample, SCAR_SAMPLE_NONE if unassigned */
    uint16_t clj;      /* consensus left junction  */
    uint16_t crj;      /* consensus right junction */
    uint16_t tlj;      /* target left junction     */
    uint16_t trj;      /* target right junction    */
    uint16_t hr_n;     /* donor occurrences in the scanned range (first -> [10][0], rest -> [10][1]) */
    int8_t   phase_r1; /* first k with phase_R1[k] == read[:n], else NONE/NA */
    int8_t   phase_r2; /* first k with phase_R2[k] == rcomp(read[-n:]), else NONE/NA */
} scar_record;         /* 16 bytes */

/* per-sample counters, one row of SCAR_N_COUNTERS int64 per sample
 * (summary_data, INDEL_Processing.py:94,156-167) */
enum {
    SCAR_CT_ASSIGNED = 0,   /* read_count_dict[index] (INDEL_Processing.py:1177) */
    SCAR_CT_PASSING,        /* summary[1]  */
    SCAR_CT_LEFT_DEL,       /* summary[2]  */
    SCAR_CT_RIGHT_DEL,      /* summary[3]  */
    SCAR_CT_INSERTION,      /* summary[4]  */
    SCAR_CT_MICROHOM,       /* summary[5]  */
    SCAR_CT_NO_JUNCTION,    /* summary[6][0] */
    SCAR_CT_NO_CUT,         /* summary[6][1] */
    SCAR_CT_FILTERED,       /* summary[7][0] */
    SCAR_CT_HR_FIRST,       /* summary[10][0] */
    SCAR_CT_HR_MORE,        /* summary[10][1] */
    SCAR_CT_N_INSERTION,    /* dropped with no reference counter (SlidingWindow.pyx:140-141) */
    SCAR_CT_SCAR,           /* len(read_results_list) */
    SCAR_N_COUNTERS = 16
};

/* one aggregated scar pattern (results_freq_dict entry, INDEL_Processing.py:186-196) */
typedef struct scar_entry {
    uint64_t h1, h2;     /* 128-bit identity of the key (sample, tlj, trj, ins|mh bytes, hr): two
                            independent 64-bit hashes; both must agree for two reads to share an entry */
    uint64_t first_idx;  /* global index of the FIRST read with this key (:193-196) */
    uint32_t count;      /* results_freq_dict[key][0] */
    uint16_t sample;
    uint16_t pad;
} scar_entry;            /* 32 bytes */

#ifdef __cplusplus
}
#endif
#endif
