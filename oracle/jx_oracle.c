/* oracle/jx_oracle.c — TEST INFRASTRUCTURE ONLY (see jx_oracle.h for the rules).
 *
 * CPU restatement, in plain single-threaded C, of the reference hot path:
 *   junction logic   /root/reference/src/junctions/junctions_extractor.cc
 *   BED12 / sort     /root/reference/src/junctions/junctions_extractor.h
 *   BAM/BGZF/BAI     /root/reference/src/utils/htslib/{sam.c,hts.c,bgzf.c}  (htslib 1.2.1)
 * Every function cites the reference lines it follows.  It is deliberately simple and
 * literal (state machine, not the closed form the CUDA path uses) so that the two
 * implementations are independent.
 */
#define _GNU_SOURCE
#include "jx_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#include <zlib.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <fcntl.h>
#include <unistd.h>

/* ------------------------------------------------------------------ table */
typedef struct {
    int32_t  tid;
    uint32_t start, end, ts, te, count, name_index;
    uint8_t  strand, proxy, l_ok, r_ok;
    int32_t  next; /* hash chain */
    /* -b: Junction::barcodes (junctions_extractor.h:57-58) as (dictionary id, count) pairs in FIRST-SEEN order */
    uint32_t* bcs; uint32_t n_bcs, cap_bcs;
} entry_t;

struct jxo {
    uint32_t min_anchor, min_intron, max_intron;
    int strandness;
    char tag[2];
    entry_t* e; size_t n, cap;
    int32_t* bucket; size_t nbucket;
    char** contig; int32_t n_contig;
    int record; jxo_candidate* cand; size_t n_cand, cap_cand;
    uint64_t reads_seen;
    char errbuf[256];
    /* FASTA (intron-motif strand inference, junctions_extractor.cc:325-359,548-584) */
    int has_fasta, fasta_error;
    int32_t n_seq; char** seq_name; uint8_t** seq; int64_t* seq_len;
    /* -b single-cell barcodes (junctions_extractor.cc:203-215,362-374) */
    int bc_on; char bc_tag[2];
    char** bc_name; uint32_t n_bc, cap_bc;
    int32_t* bc_hash; uint32_t n_bc_hash;  /* open addressing over bc_name (FNV-1a), -1 = empty */
    uint32_t cur_bc;                      /* barcode of the alignment being fed (j1.barcodes, one entry) */
    uint64_t bc_missing;
};

jxo_t* jxo_new(uint32_t min_anchor, uint32_t min_intron, uint32_t max_intron,
               int strandness, const char* strand_tag) {
    jxo_t* o = (jxo_t*)calloc(1, sizeof(*o));
    o->min_anchor = min_anchor; o->min_intron = min_intron; o->max_intron = max_intron;
    o->strandness = strandness;
    o->tag[0] = strand_tag && strand_tag[0] ? strand_tag[0] : 'X';
    o->tag[1] = strand_tag && strand_tag[0] ? strand_tag[1] : 'S';
    o->nbucket = 1u << 16;
    o->bucket = (int32_t*)malloc(o->nbucket * sizeof(int32_t));
    memset(o->bucket, 0xff, o->nbucket * sizeof(int32_t));
    return o;
}

void jxo_free(jxo_t* o) {
    if (!o) return;
    for (int32_t i = 0; i < o->n_contig; ++i) free(o->contig[i]);
    for (size_t i = 0; i < o->n; ++i) free(o->e[i].bcs);
    for (uint32_t i = 0; i < o->n_bc; ++i) free(o->bc_name[i]);
    free(o->bc_name); free(o->bc_hash);
    free(o->contig); free(o->e); free(o->bucket); free(o->cand);
    for (int32_t i = 0; i < o->n_seq; ++i) { free(o->seq_name[i]); free(o->seq[i]); }
    free(o->seq_name); free(o->seq); free(o->seq_len);
    free(o);
}

/* FASTA as faidx sees it (htslib faidx.c:82-155 fai_build_core): a sequence's name is the header up to the first
 * white space, its bases are the isgraph() characters of the following lines, duplicates of a name are ignored.
 * Returns 0, or -1 if the file cannot be read. */
int jxo_set_fasta(jxo_t* o, const char* path) {
    FILE* f = fopen(path, "rb");
    if (!f) return -1;
    fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
    char* buf = (char*)malloc((size_t)sz + 1);
    if (fread(buf, 1, (size_t)sz, f) != (size_t)sz) { fclose(f); free(buf); return -1; }
    fclose(f); buf[sz] = 0;
    long i = 0;
    while (i < sz) {
        if (buf[i] != '>') { ++i; continue; }
        long j = i + 1;
        while (j < sz && buf[j] != '\n' && (buf[j] == ' ' || buf[j] == '\t')) ++j;    /* leading blanks before the name */
        long n0 = j;
        while (j < sz && !(buf[j] == ' ' || buf[j] == '\t' || buf[j] == '\n' || buf[j] == '\r' || buf[j] == '\v' || buf[j] == '\f')) ++j;
        char* name = strndup(buf + n0, (size_t)(j - n0));
        while (j < sz && buf[j] != '\n') ++j;
        long k = j;
        while (k < sz && buf[k] != '>') ++k;            /* sequence lines run to the next header (a '>' inside a line is not handled) */
        int dup = 0;
        for (int32_t q = 0; q < o->n_seq; ++q) if (!strcmp(o->seq_name[q], name)) dup = 1;
        if (!dup) {
            uint8_t* sq = (uint8_t*)malloc((size_t)(k - j) + 1);
            int64_t l = 0;
            for (long q = j; q < k; ++q) if ((unsigned char)buf[q] > 32 && (unsigned char)buf[q] < 127) sq[l++] = (uint8_t)buf[q];
            o->seq_name = (char**)realloc(o->seq_name, sizeof(char*) * (size_t)(o->n_seq + 1));
            o->seq = (uint8_t**)realloc(o->seq, sizeof(uint8_t*) * (size_t)(o->n_seq + 1));
            o->seq_len = (int64_t*)realloc(o->seq_len, sizeof(int64_t) * (size_t)(o->n_seq + 1));
            o->seq_name[o->n_seq] = name; o->seq[o->n_seq] = sq; o->seq_len[o->n_seq] = l; o->n_seq++;
        } else free(name);
        i = k;
    }
    free(buf);
    o->has_fasta = 1;
    return 0;
}
const char* jxo_error(const jxo_t* o) { return o->fasta_error ? o->errbuf : NULL; }

/* fai_fetch (faidx.c:341-415) for "chrom:b-e" with 1-based inclusive b, e: bases [b-1, e) clipped to the sequence. */
static int fetch2(const jxo_t* o, int32_t q, int64_t b1, int64_t e1, uint8_t out[2]) {
    int64_t len = o->seq_len[q], beg = b1, end = e1;
    if (beg > 0) --beg;
    if (beg >= len) beg = len;
    if (end >= len) end = len;
    if (beg > end) beg = end;
    int n = 0;
    for (int64_t x = beg; x < end && n < 2; ++x) out[n++] = o->seq[q][x];
    return (int)(end - beg);
}
static uint8_t comp_base(uint8_t c) {                    /* common.h:59-83 rev_comp */
    switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; default: return 'N'; }
}
/* get_splice_site (:564-584) + set_junction_strand_intron_motif (:325-342).  prev = strand the Junction object
 * carries from the previous junction of the same read (0 for the first).  Returns '+', '-' or '?'. */
static uint8_t motif_strand(jxo_t* o, int32_t tid, uint32_t start, uint32_t end, uint8_t prev) {
    const char* chrom = jxo_contig(o, tid);
    int32_t q = -1;
    for (int32_t i = 0; i < o->n_seq; ++i) if (!strcmp(o->seq_name[i], chrom)) { q = i; break; }
    if (q < 0) {                                          /* fai_fetch returns NULL -> runtime_error (:553-555) */
        if (!o->fasta_error) {
            o->fasta_error = 1;
            snprintf(o->errbuf, sizeof o->errbuf, "Unable to extract FASTA sequence for position %s:%u-%u\n\n", chrom,
                     start + 1u, start + 2u);
        }
        return '?';
    }
    uint8_t s1[2] = {0, 0}, s2[2] = {0, 0};
    int l1 = fetch2(o, q, (int64_t)(uint32_t)(start + 1u), (int64_t)(uint32_t)(start + 2u), s1);
    int l2 = fetch2(o, q, (int64_t)(uint32_t)(end + 1u - 2u), (int64_t)(uint32_t)(end + 1u - 1u), s2);
    if (l1 != 2 || l2 != 2) return '?';                   /* a shorter string cannot equal any 5-character motif */
    uint8_t m[4];
    if (prev == '-') { m[0] = comp_base(s2[1]); m[1] = comp_base(s2[0]); m[2] = comp_base(s1[1]); m[3] = comp_base(s1[0]); }
    else { m[0] = s1[0]; m[1] = s1[1]; m[2] = s2[0]; m[3] = s2[1]; }
    static const char* plus[3] = {"GTAG", "GCAG", "ATAC"};
    static const char* minus[3] = {"CTAC", "CTGC", "GTAT"};
    for (int i = 0; i < 3; ++i) if (!memcmp(m, plus[i], 4)) return '+';
    for (int i = 0; i < 3; ++i) if (!memcmp(m, minus[i], 4)) return '-';
    return '?';
}

void jxo_set_contigs(jxo_t* o, int32_t n, const char* const* names) {
    for (int32_t i = 0; i < o->n_contig; ++i) free(o->contig[i]);
    free(o->contig);
    o->contig = (char**)calloc(n > 0 ? n : 1, sizeof(char*));
    o->n_contig = n;
    for (int32_t i = 0; i < n; ++i) o->contig[i] = strdup(names[i]);
}

const char* jxo_contig(const jxo_t* o, int32_t tid) {
    static char buf[32];
    if (tid >= 0 && tid < o->n_contig) return o->contig[tid];
    snprintf(buf, sizeof buf, "tid%d", tid);
    return buf;
}

static uint64_t key_hash(int32_t tid, uint32_t s, uint32_t e, uint8_t proxy) {
    uint64_t h = ((uint64_t)s << 32 | e) * 0x9E3779B97F4A7C15ull;
    h ^= ((uint64_t)(uint32_t)tid << 2 | proxy) * 0xC2B2AE3D27D4EB4Full;
    h ^= h >> 29;
    return h;
}

static void rehash(jxo_t* o) {
    free(o->bucket);
    o->nbucket *= 4;
    o->bucket = (int32_t*)malloc(o->nbucket * sizeof(int32_t));
    memset(o->bucket, 0xff, o->nbucket * sizeof(int32_t));
    for (size_t i = 0; i < o->n; ++i) {
        entry_t* x = &o->e[i];
        size_t b = key_hash(x->tid, x->start, x->end, x->proxy) & (o->nbucket - 1);
        x->next = o->bucket[b];
        o->bucket[b] = (int32_t)i;
    }
}

/* -b mode: what set_junction_barcode (junctions_extractor.cc:362-374) puts into j1.barcodes for one alignment.
 * The dictionary is an array of names with a small open-addressing index. */
void jxo_enable_barcodes(jxo_t* o, const char* tag) {
    o->bc_on = 1;
    o->bc_tag[0] = tag && tag[0] ? tag[0] : 'C';
    o->bc_tag[1] = tag && tag[0] ? tag[1] : 'B';
}
static uint32_t bc_fnv(const char* s) { uint32_t h = 2166136261u; for (; *s; ++s) { h ^= (uint8_t)*s; h *= 16777619u; } return h; }
void jxo_set_read_barcode(jxo_t* o, const char* bc) {
    if (2 * (o->n_bc + 1) > o->n_bc_hash) {                  /* grow + rebuild, load <= 0.5 */
        o->n_bc_hash = o->n_bc_hash ? o->n_bc_hash * 4 : 1024;
        free(o->bc_hash);
        o->bc_hash = (int32_t*)malloc(o->n_bc_hash * sizeof(int32_t));
        memset(o->bc_hash, 0xff, o->n_bc_hash * sizeof(int32_t));
        for (uint32_t i = 0; i < o->n_bc; ++i) {
            uint32_t h = bc_fnv(o->bc_name[i]) & (o->n_bc_hash - 1);
            while (o->bc_hash[h] >= 0) h = (h + 1) & (o->n_bc_hash - 1);
            o->bc_hash[h] = (int32_t)i;
        }
    }
    uint32_t h = bc_fnv(bc) & (o->n_bc_hash - 1);
    for (; o->bc_hash[h] >= 0; h = (h + 1) & (o->n_bc_hash - 1))
        if (!strcmp(o->bc_name[o->bc_hash[h]], bc)) { o->cur_bc = (uint32_t)o->bc_hash[h]; return; }
    if (o->n_bc == o->cap_bc) {
        o->cap_bc = o->cap_bc ? o->cap_bc * 2 : 256;
        o->bc_name = (char**)realloc(o->bc_name, o->cap_bc * sizeof(char*));
    }
    o->bc_name[o->n_bc] = strdup(bc);
    o->bc_hash[h] = (int32_t)o->n_bc;
    o->cur_bc = o->n_bc++;
}
uint64_t jxo_barcodes_missing(const jxo_t* o) { return o->bc_missing; }

/* :203-215 — the stored map is replaced by a copy of itself with this read's barcode counted (found: ++, else insert) */
static void entry_count_barcode(entry_t* x, uint32_t bc) {
    for (uint32_t i = 0; i < x->n_bcs; ++i)
        if (x->bcs[2 * i] == bc) { x->bcs[2 * i + 1] += 1; return; }
    if (x->n_bcs == x->cap_bcs) {
        x->cap_bcs = x->cap_bcs ? x->cap_bcs * 2 : 4;
        x->bcs = (uint32_t*)realloc(x->bcs, 2 * (size_t)x->cap_bcs * sizeof(uint32_t));
    }
    x->bcs[2 * x->n_bcs] = bc; x->bcs[2 * x->n_bcs + 1] = 1;
    x->n_bcs += 1;
}

/* add_junction = junction_qc + keyed merge.
 * junctions_extractor.cc:160-170 (qc), :174-235 (merge), :152-157 (name = map.size()+1). */
void jxo_add(jxo_t* o, int32_t tid, uint32_t start, uint32_t end, uint32_t ts, uint32_t te,
             uint8_t strand_char) {
    uint32_t ilen = end - start;                                  /* uint32 arithmetic, :161-162 */
    if (ilen < o->min_intron || ilen > o->max_intron) return;
    uint8_t l_ok = (uint32_t)(start - ts) >= o->min_anchor;        /* :165-168 */
    uint8_t r_ok = (uint32_t)(te - end) >= o->min_anchor;
    uint8_t proxy = strand_char == '+' ? 0 : strand_char == '-' ? 1 : 2;   /* :186-193 */
    size_t b = key_hash(tid, start, end, proxy) & (o->nbucket - 1);
    for (int32_t i = o->bucket[b]; i >= 0; i = o->e[i].next) {
        entry_t* x = &o->e[i];
        if (x->tid == tid && x->start == start && x->end == end && x->proxy == proxy) {
            x->count += 1;                                        /* :215 */
            if (ts < x->ts) x->ts = ts;                           /* :220-223 */
            if (te > x->te) x->te = te;
            x->l_ok |= l_ok; x->r_ok |= r_ok;                     /* :225-226 */
            x->strand = strand_char;                              /* :229 stored object := new read's copy */
            if (o->bc_on) entry_count_barcode(x, o->cur_bc);      /* :204-216 */
            return;
        }
    }
    if (o->n == o->cap) {
        o->cap = o->cap ? o->cap * 2 : 1024;
        o->e = (entry_t*)realloc(o->e, o->cap * sizeof(entry_t));
    }
    entry_t* x = &o->e[o->n];
    x->tid = tid; x->start = start; x->end = end; x->ts = ts; x->te = te;
    x->count = 1; x->name_index = (uint32_t)(o->n + 1);           /* :152-157, :198 */
    x->strand = strand_char; x->proxy = proxy; x->l_ok = l_ok; x->r_ok = r_ok;
    x->bcs = NULL; x->n_bcs = x->cap_bcs = 0;
    if (o->bc_on) entry_count_barcode(x, o->cur_bc);              /* new junction: j1.barcodes = {barcode: 1} (:394,:229) */
    x->next = o->bucket[b];
    o->bucket[b] = (int32_t)o->n;
    o->n += 1;
    if (o->n > o->nbucket * 2) rehash(o);
}

/* ------------------------------------------------------------------ per read */
/* set_junction_strand without FASTA: junctions_extractor.cc:345-359 ->
 *   _XS   :283-294 (bam_aux2A: 0 unless type 'A'; 0 -> '?')
 *   _flag :297-322 */
static uint8_t read_strand(const jxo_t* o, uint32_t flag, uint8_t strand_byte) {
    if (o->strandness == 0) return strand_byte ? strand_byte : '?';
    int reversed = (flag >> 4) % 2, mate_reversed = (flag >> 5) % 2;
    int first_in_pair = (flag >> 6) % 2, second_in_pair = (flag >> 7) % 2;
    int bool_strandness = o->strandness - 1;
    int first_strand = !bool_strandness ^ first_in_pair ^ reversed;
    int second_strand = !bool_strandness ^ second_in_pair ^ mate_reversed;
    if (first_strand != second_strand) return '?';
    return first_strand ? '+' : '-';
}

/* `strand` is the read's XS/flag strand; with a FASTA the motif decides first and the read's strand is only the
 * fall-back for '?' (set_junction_strand :345-359).  *jstrand is the strand field of the reference's reused
 * Junction object j1. */
static void emit(jxo_t* o, int32_t tid, uint32_t start, uint32_t end, uint32_t ts, uint32_t te,
                 uint8_t strand, uint64_t read_index, uint32_t k, uint8_t* jstrand) {
    if (o->fasta_error) return;                           /* the reference has thrown: nothing else is added */
    if (o->has_fasta) {
        uint8_t m = motif_strand(o, tid, start, end, *jstrand);
        if (o->fasta_error) return;
        if (m != '?') strand = m;
    }
    *jstrand = strand;
    if (o->record) {
        if (o->n_cand == o->cap_cand) {
            o->cap_cand = o->cap_cand ? o->cap_cand * 2 : 1024;
            o->cand = (jxo_candidate*)realloc(o->cand, o->cap_cand * sizeof(jxo_candidate));
        }
        jxo_candidate* c = &o->cand[o->n_cand++];
        memset(c, 0, sizeof *c);
        c->start = start; c->end = end; c->thick_start = ts; c->thick_end = te;
        c->read_index = read_index; c->tid = tid; c->k = (uint16_t)k; c->strand = strand;
    }
    jxo_add(o, tid, start, end, ts, te, strand);
}

/* parse_alignment_into_junctions, junctions_extractor.cc:377-497, literally. */
void jxo_read(jxo_t* o, int32_t tid, int32_t pos, uint32_t flag, uint8_t strand_byte,
              const uint32_t* cigar, uint32_t n_cigar) {
    uint64_t ridx = o->reads_seen++;
    if (n_cigar <= 1) return;                                     /* :379 */
    /* :383-384 `string chr(header->target_name[chr_id])` is undefined behaviour for a tid outside
     * the header (the reference crashes); both the oracle and the product skip such reads. */
    if (tid < 0) return;
    uint32_t start = (uint32_t)pos, thick_start = (uint32_t)pos, end = 0, thick_end = 0;
    int started = 0;
    uint32_t open_k = 0;  /* index of the N op that opened the pending junction */
    uint8_t strand = read_strand(o, flag, strand_byte);
    uint8_t jstrand = 0;                                          /* j1.strand == "" before the first junction */
    for (uint32_t i = 0; i < n_cigar; ++i) {
        uint32_t op = cigar[i] & 0xf, len = cigar[i] >> 4;         /* htslib/sam.h:75-83 */
        switch (op) {
        case 3: /* N :403-431 */
            if (!started) {
                end = start + len; thick_end = end; started = 1; open_k = i;
            } else {
                emit(o, tid, start, end, thick_start, thick_end, strand, ridx, open_k, &jstrand);
                thick_start = end; start = thick_end; end = start + len; thick_end = end;
                started = 1; open_k = i;
            }
            break;
        case 0: case 7: /* M, = :432-438 */
            if (!started) start += len; else thick_end += len;
            break;
        case 2: case 8: /* D, X :440-459 */
            if (!started) { start += len; thick_start = start; }
            else {
                emit(o, tid, start, end, thick_start, thick_end, strand, ridx, open_k, &jstrand);
                start = thick_end + len; thick_start = start;
            }
            started = 0;
            break;
        case 1: case 4: /* I, S :460-478 */
            if (!started) thick_start = start;
            else {
                emit(o, tid, start, end, thick_start, thick_end, strand, ridx, open_k, &jstrand);
                start = thick_end; thick_start = start;
            }
            started = 0;
            break;
        case 5: /* H :479-480 */
            break;
        default: /* P, B, 10-15: "Unknown cigar" on stderr, no state change :481-483 */
            break;
        }
    }
    if (started)                                                  /* :485-495 */
        emit(o, tid, start, end, thick_start, thick_end, strand, ridx, open_k, &jstrand);
}

void jxo_batch(jxo_t* o, uint32_t n_reads, const int32_t* tid, const int32_t* pos,
               const uint32_t* meta, const uint32_t* cig_off, const uint32_t* cigar) {
    for (uint32_t i = 0; i < n_reads; ++i)
        jxo_read(o, tid[i], pos[i], meta[i] >> 16, (uint8_t)(meta[i] & 0xff),
                 cigar + cig_off[i], cig_off[i + 1] - cig_off[i]);
}

/* -b over a SoA batch: names[bc[i]] is what set_junction_barcode (:362-374) would have found for alignment i (read for
 * n_cigar > 1 alignments only, like the reference's call at :393-395). */
void jxo_batch_barcodes(jxo_t* o, uint32_t n_reads, const int32_t* tid, const int32_t* pos, const uint32_t* meta,
                        const uint32_t* cig_off, const uint32_t* cigar, const uint32_t* bc, const char* const* names) {
    for (uint32_t i = 0; i < n_reads; ++i) {
        const uint32_t n = cig_off[i + 1] - cig_off[i];
        if (n > 1) jxo_set_read_barcode(o, names[bc[i]]);
        jxo_read(o, tid[i], pos[i], meta[i] >> 16, (uint8_t)(meta[i] & 0xff), cigar + cig_off[i], n);
    }
}

void jxo_record_candidates(jxo_t* o, int on) { o->record = on; }
size_t jxo_candidates(const jxo_t* o, jxo_candidate* out, size_t cap) {
    size_t n = o->n_cand < cap ? o->n_cand : cap;
    if (out && n) memcpy(out, o->cand, n * sizeof(jxo_candidate));
    return o->n_cand;
}
uint64_t jxo_reads_seen(const jxo_t* o) { return o->reads_seen; }

/* ------------------------------------------------------------------ sort + print */
static const jxo_t* g_sort_ctx;
/* compare_junctions, junctions_extractor.h:117-140: chrom (std::string <), thick_start,
 * thick_end, name (std::string < on "JUNC%08d": wider than 8 digits sorts as a string). */
static int cmp_entry(const void* pa, const void* pb) {
    const entry_t* a = (const entry_t*)pa; const entry_t* b = (const entry_t*)pb;
    if (a->tid != b->tid) {
        int c = strcmp(jxo_contig(g_sort_ctx, a->tid), jxo_contig(g_sort_ctx, b->tid));
        if (c) return c;
    }
    if (a->ts != b->ts) return a->ts < b->ts ? -1 : 1;
    if (a->te != b->te) return a->te < b->te ? -1 : 1;
    char na[32], nb[32];
    snprintf(na, sizeof na, "JUNC%08d", (int)a->name_index);
    snprintf(nb, sizeof nb, "JUNC%08d", (int)b->name_index);
    return strcmp(na, nb);
}

static entry_t* sorted_copy(jxo_t* o) {
    entry_t* v = (entry_t*)malloc((o->n ? o->n : 1) * sizeof(entry_t));
    memcpy(v, o->e, o->n * sizeof(entry_t));
    /* contig names may compare equal for different tids (duplicate @SQ); jxo_contig's static
     * buffer is only used for out-of-range tids, which strcmp would alias: resolve first. */
    g_sort_ctx = o;
    qsort(v, o->n, sizeof(entry_t), cmp_entry);
    return v;
}

size_t jxo_count(const jxo_t* o) { return o->n; }

size_t jxo_get(jxo_t* o, jxo_junction* out, size_t cap) {
    entry_t* v = sorted_copy(o);
    size_t n = o->n < cap ? o->n : cap;
    for (size_t i = 0; i < n; ++i) {
        jxo_junction* j = &out[i];
        j->tid = v[i].tid; j->start = v[i].start; j->end = v[i].end;
        j->thick_start = v[i].ts; j->thick_end = v[i].te; j->read_count = v[i].count;
        j->name_index = v[i].name_index; j->strand = v[i].strand;
        j->left_ok = v[i].l_ok; j->right_ok = v[i].r_ok; j->pad = 0;
    }
    free(v);
    return o->n;
}

/* Junction::print, junctions_extractor.h:90-98; filter :267. */
int jxo_write_bed12(jxo_t* o, FILE* out) {
    entry_t* v = sorted_copy(o);
    for (size_t i = 0; i < o->n; ++i) {
        const entry_t* x = &v[i];
        if (!(x->l_ok && x->r_ok)) continue;
        fprintf(out, "%s\t%u\t%u\tJUNC%08d\t%u\t%c\t%u\t%u\t255,0,0\t2\t%u,%u\t0,%u\n",
                jxo_contig(o, x->tid), x->ts, x->te, (int)x->name_index, x->count, x->strand,
                x->ts, x->te, (uint32_t)(x->start - x->ts), (uint32_t)(x->te - x->end),
                (uint32_t)(x->end - x->ts));
    }
    free(v);
    return 0;
}

/* The -b file (print_barcodes, junctions_extractor.h:99-111, for the junctions :267-273 prints, in that order) lists
 * each junction's unordered_map in libstdc++ iteration order.  That order is a property of the library; the map went
 * through one copy-assignment per read (copies keep bucket count and node order), so it equals the order of ONE map
 * filled with the junction's distinct barcodes in first-seen order.  This writes that replay input, one line per printed
 * junction: "bc count bc count ..."; oracle/bc_replay.cc (real std::unordered_map) turns it into the file's bytes. */
int jxo_write_barcode_replay_path(jxo_t* o, const char* path) {
    FILE* f = fopen(path, "w");
    if (!f) return -1;
    entry_t* v = sorted_copy(o);
    for (size_t i = 0; i < o->n; ++i) {
        const entry_t* x = &v[i];
        if (!(x->l_ok && x->r_ok)) continue;
        for (uint32_t k = 0; k < x->n_bcs; ++k)
            fprintf(f, "%s%s %u", k ? " " : "", o->bc_name[x->bcs[2 * k]], x->bcs[2 * k + 1]);
        fputc('\n', f);
    }
    free(v);
    return fclose(f);
}

int jxo_write_bed12_path(jxo_t* o, const char* path) {
    FILE* f = fopen(path, "w");
    if (!f) return -1;
    jxo_write_bed12(o, f);
    return fclose(f);
}

/* ------------------------------------------------------------------ BGZF */
typedef struct {
    const uint8_t* file; size_t size;     /* mmapped */
    uint64_t fpos;                        /* htell(): compressed offset of the next block to read */
    uint64_t block_address;
    int32_t  block_length, block_offset;
    uint8_t  buf[0x10000];
} bgzf_t;

/* bgzf_read_block, bgzf.c:421-546 + inflate_block :292-316: 18-byte header (check_header
 * :348-355), BSIZE at +16, raw deflate, CRC32 + ISIZE trailer (CRC not verified by the
 * reference).  Returns 0 with block_length==0 at end of file, -1 on a malformed block. */
static int bgzf_read_block_(bgzf_t* b) {
    uint64_t addr = b->fpos;
    if (addr >= b->size) { b->block_length = 0; return 0; }
    if (addr + 18 > b->size) return -1;
    const uint8_t* h = b->file + addr;
    if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8) return -1;
    uint32_t xlen = h[10] | h[11] << 8;
    if (!(h[3] & 4) || xlen != 6 || h[12] != 'B' || h[13] != 'C' || h[14] != 2 || h[15] != 0)
        return -1;   /* plain gzip is accepted by the reference but cannot be indexed */
    uint32_t bsize = (h[16] | h[17] << 8) + 1u;
    if (addr + bsize > b->size || bsize < 26) return -1;
    z_stream zs; memset(&zs, 0, sizeof zs);
    /* bgzf.c:298-299 passes block_length - 16 bytes: deflate data + the 8-byte trailer + 2 bytes of whatever the buffer held
     * behind the block (not reproducible); the trailer is offered here, a damaged stream may run on into it */
    zs.next_in = (Bytef*)(h + 18); zs.avail_in = bsize - 18;
    zs.next_out = b->buf; zs.avail_out = sizeof b->buf;
    if (inflateInit2(&zs, -15) != Z_OK) return -1;
    int rc = inflate(&zs, Z_FINISH);
    inflateEnd(&zs);
    if (rc != Z_STREAM_END) return -1;
    b->fpos = addr + bsize;
    if (b->block_length != 0) b->block_offset = 0;   /* keep the offset right after a seek */
    b->block_address = addr;
    b->block_length = (int32_t)zs.total_out;
    return 0;
}

/* bgzf_read, bgzf.c:548-577.  Note the `break` on an EMPTY block: an ISIZE=0 block in the
 * middle of a file yields a short read, which bam_read1 treats as end of file. */
static long bgzf_read_(bgzf_t* b, void* dst, size_t n) {
    size_t got = 0;
    while (got < n) {
        int32_t available = b->block_length - b->block_offset;
        if (available <= 0) {
            if (bgzf_read_block_(b) != 0) return -1;
            available = b->block_length - b->block_offset;
            if (available <= 0) break;
        }
        size_t take = n - got < (size_t)available ? n - got : (size_t)available;
        memcpy((uint8_t*)dst + got, b->buf + b->block_offset, take);
        b->block_offset += (int32_t)take; got += take;
    }
    if (b->block_offset == b->block_length) {
        b->block_address = b->fpos;
        b->block_offset = b->block_length = 0;
    }
    return (long)got;
}

/* bgzf_seek, bgzf.c:848-867 */
static void bgzf_seek_(bgzf_t* b, uint64_t voff) {
    b->fpos = voff >> 16;
    b->block_length = 0;
    b->block_address = voff >> 16;
    b->block_offset = (int32_t)(voff & 0xffff);
}

/* bgzf_tell, htslib/bgzf.h: (block_address << 16) | (block_offset & 0xFFFF) */
static uint64_t bgzf_tell_(bgzf_t* b) {
    return b->block_address << 16 | ((uint64_t)b->block_offset & 0xffff);
}

/* ------------------------------------------------------------------ BAI */
typedef struct { uint64_t u, v; } pair64;
typedef struct { uint32_t bin; uint64_t loff; int32_t n; pair64* list; } bin_t;
typedef struct { int32_t n_bin; bin_t* bins; int32_t n_intv; uint64_t* ioff; } refidx_t;
typedef struct { int32_t n_ref; refidx_t* ref; uint64_t n_no_coor; } bai_t;
#define META_BIN 37450u   /* hts.c:1092 with min_shift 14, n_lvls 5 */

static void bai_free(bai_t* x) {
    if (!x) return;
    for (int32_t i = 0; i < x->n_ref; ++i) {
        for (int32_t j = 0; j < x->ref[i].n_bin; ++j) free(x->ref[i].bins[j].list);
        free(x->ref[i].bins); free(x->ref[i].ioff);
    }
    free(x->ref); free(x);
}

static bin_t* find_bin(refidx_t* r, uint32_t bin) {
    for (int32_t j = 0; j < r->n_bin; ++j) if (r->bins[j].bin == bin) return &r->bins[j];
    return NULL;
}

/* hts_idx_load_local / hts_idx_load_core, hts.c:1569-1624,1517-1567; update_loff hts.c:1193-1222 */
static bai_t* bai_load(const char* fn) {
    FILE* f = fopen(fn, "rb");
    if (!f) return NULL;
    char magic[4];
    bai_t* x = (bai_t*)calloc(1, sizeof *x);
    if (fread(magic, 1, 4, f) != 4 || memcmp(magic, "BAI\1", 4)) goto fail;
    if (fread(&x->n_ref, 4, 1, f) != 1 || x->n_ref < 0) goto fail;
    x->ref = (refidx_t*)calloc(x->n_ref ? x->n_ref : 1, sizeof(refidx_t));
    for (int32_t i = 0; i < x->n_ref; ++i) {
        refidx_t* r = &x->ref[i];
        if (fread(&r->n_bin, 4, 1, f) != 1) goto fail;
        r->bins = (bin_t*)calloc(r->n_bin ? r->n_bin : 1, sizeof(bin_t));
        for (int32_t j = 0; j < r->n_bin; ++j) {
            bin_t* b = &r->bins[j];
            if (fread(&b->bin, 4, 1, f) != 1 || fread(&b->n, 4, 1, f) != 1) goto fail;
            b->list = (pair64*)malloc((b->n ? b->n : 1) * sizeof(pair64));
            if (fread(b->list, 16, b->n, f) != (size_t)b->n) goto fail;
        }
        if (fread(&r->n_intv, 4, 1, f) != 1) goto fail;
        r->ioff = (uint64_t*)malloc((r->n_intv ? r->n_intv : 1) * 8);
        if (fread(r->ioff, 8, r->n_intv, f) != (size_t)r->n_intv) goto fail;
        for (int32_t j = 1; j < r->n_intv; ++j)                     /* hts.c:1559-1560 */
            if (r->ioff[j] == 0) r->ioff[j] = r->ioff[j - 1];
        for (int32_t j = 0; j < r->n_bin; ++j) {                    /* update_loff */
            bin_t* b = &r->bins[j];
            if (b->bin < 37449u) {
                /* hts_bin_bot: first leaf-level bin under b, minus first leaf id */
                uint32_t bin = b->bin; int l = 0;
                for (uint32_t t = bin; t; t = (t - 1) >> 3) ++l;
                uint32_t first_of_level = ((1u << (3 * l)) - 1) / 7;
                int64_t bot = (int64_t)(bin - first_of_level) << (3 * (5 - l));
                b->loff = bot < r->n_intv ? r->ioff[bot] : 0;
            } else b->loff = 0;
        }
    }
    if (fread(&x->n_no_coor, 8, 1, f) != 1) x->n_no_coor = 0;
    fclose(f);
    return x;
fail:
    fclose(f); bai_free(x);
    return NULL;
}

/* hts_idx_getfn, hts.c:2009-2029: "<bam>.bai" then "<stem>.bai" (stem = up to last '.').
 * The reference tries .csi first (hts.c:2031-2042); this oracle only reads .bai. */
static bai_t* bai_find(const char* bam) {
    size_t l = strlen(bam);
    char* fn = (char*)malloc(l + 8);
    sprintf(fn, "%s.bai", bam);
    bai_t* x = bai_load(fn);
    if (!x) {
        size_t i;
        for (i = l - 1; i > 0; --i) if (bam[i] == '.') break;
        memcpy(fn, bam, i); strcpy(fn + i, ".bai");
        x = bai_load(fn);
    }
    free(fn);
    return x;
}

/* ------------------------------------------------------------------ BAM */
typedef struct {
    int32_t tid, pos; uint32_t flag, n_cigar, l_qname; int32_t l_qseq;
    uint8_t* data; int32_t l_data, m_data;
} rec_t;

/* bam_read1, sam.c:399-432.  <0 on EOF / malformed (the caller's loop just ends, :525). */
static int read_rec(bgzf_t* fp, rec_t* r) {
    int32_t block_len; uint32_t x[8];
    if (bgzf_read_(fp, &block_len, 4) != 4) return -1;
    if (bgzf_read_(fp, x, 32) != 32) return -3;
    r->tid = (int32_t)x[0]; r->pos = (int32_t)x[1];
    r->l_qname = x[2] & 0xff; r->flag = x[3] >> 16; r->n_cigar = x[3] & 0xffff;
    r->l_qseq = (int32_t)x[4];
    r->l_data = block_len - 32;
    if (r->l_data < 0 || r->l_qseq < 0 || r->l_qname < 1) return -4;
    int64_t aux_off = (int64_t)r->l_qname + 4ll * r->n_cigar + ((int64_t)r->l_qseq + 1) / 2 + r->l_qseq;
    if (aux_off > r->l_data) return -4;
    if (r->m_data < r->l_data) {
        r->m_data = r->l_data + 64;
        r->data = (uint8_t*)realloc(r->data, r->m_data);
    }
    if (bgzf_read_(fp, r->data, r->l_data) != (long)r->l_data) return -4;
    return 4 + block_len;
}

/* bam_endpos, sam.c:336-342 (cigar type bit 2 = consumes reference: M D N = X). */
static int32_t rec_endpos(const rec_t* r) {
    if (!(r->flag & 4) && r->n_cigar > 0) {
        const uint32_t* c = (const uint32_t*)(r->data + r->l_qname);
        int32_t l = 0;
        for (uint32_t k = 0; k < r->n_cigar; ++k) {
            uint32_t op = c[k] & 0xf;
            if ((0x3C1A7 >> (op << 1) & 3) & 2) l += (int32_t)(c[k] >> 4);
        }
        return r->pos + l;
    }
    return r->pos + 1;
}

/* bam_aux_get + bam_aux2A, sam.c:1254-1266,1301-1307,1233-1252. */
static uint8_t rec_strand_byte(const rec_t* r, const char tag[2]) {
    const uint8_t* s = r->data + r->l_qname + 4 * r->n_cigar + (r->l_qseq + 1) / 2 + r->l_qseq;
    const uint8_t* e = r->data + r->l_data;
    while (s + 3 <= e) {
        int match = s[0] == (uint8_t)tag[0] && s[1] == (uint8_t)tag[1];
        uint8_t type = s[2];
        s += 3;
        if (match) return type == 'A' && s < e ? *s : 0;
        switch (type) {
        case 'A': case 'c': case 'C': s += 1; break;
        case 's': case 'S': s += 2; break;
        case 'i': case 'I': case 'f': s += 4; break;
        case 'd': s += 8; break;
        case 'Z': case 'H': while (s < e && *s) ++s; ++s; break;
        case 'B': {
            if (s + 5 > e) return 0;
            uint8_t sub = *s++; uint32_t n; memcpy(&n, s, 4); s += 4;
            uint32_t sz = (sub == 'c' || sub == 'C' || sub == 'A') ? 1 : (sub == 's' || sub == 'S') ? 2 :
                          (sub == 'i' || sub == 'I' || sub == 'f') ? 4 : sub == 'd' ? 8 : 0;
            s += (size_t)sz * n; break; }
        default: return 0;   /* the reference abort()s here (sam.c:1246-1247) */
        }
    }
    return 0;
}

/* bam_aux_get + bam_aux2Z (sam.c:1254-1266,1309-1315) for the barcode tag: 1 and *val = the string, 0 = absent.
 * A tag of another type makes the reference construct std::string(NULL) and die; the oracle treats it as absent. */
static int rec_barcode(const rec_t* r, const char tag[2], const char** val) {
    const uint8_t* s = r->data + r->l_qname + 4 * r->n_cigar + (r->l_qseq + 1) / 2 + r->l_qseq;
    const uint8_t* e = r->data + r->l_data;
    while (s + 3 <= e) {
        int match = s[0] == (uint8_t)tag[0] && s[1] == (uint8_t)tag[1];
        uint8_t type = s[2];
        s += 3;
        if (match) { if (type != 'Z' && type != 'H') return 0; *val = (const char*)s; return 1; }
        switch (type) {
        case 'A': case 'c': case 'C': s += 1; break;
        case 's': case 'S': s += 2; break;
        case 'i': case 'I': case 'f': s += 4; break;
        case 'd': s += 8; break;
        case 'Z': case 'H': while (s < e && *s) ++s; ++s; break;
        case 'B': {
            if (s + 5 > e) return 0;
            uint8_t sub = *s++; uint32_t n; memcpy(&n, s, 4); s += 4;
            uint32_t sz = (sub == 'c' || sub == 'C' || sub == 'A') ? 1 : (sub == 's' || sub == 'S') ? 2 :
                          (sub == 'i' || sub == 'I' || sub == 'f') ? 4 : sub == 'd' ? 8 : 0;
            s += (size_t)sz * n; break; }
        default: return 0;
        }
    }
    return 0;
}

static void feed(jxo_t* o, const rec_t* r) {
    if (o->bc_on && r->n_cigar > 1) {                      /* set_junction_barcode, junctions_extractor.cc:393-395,362-374 */
        const char* bc = NULL;
        if (rec_barcode(r, o->bc_tag, &bc)) jxo_set_read_barcode(o, bc);
        else { jxo_set_read_barcode(o, "?"); o->bc_missing += 1; }
    }
    uint8_t sb = 0;
    if (o->strandness == 0 && r->n_cigar > 1) sb = rec_strand_byte(r, o->tag);
    jxo_read(o, r->tid >= o->n_contig ? -1 : r->tid, r->pos, r->flag, sb, (const uint32_t*)(r->data + r->l_qname), r->n_cigar);
}

/* hts_parse_decimal with HTS_PARSE_THOUSANDS_SEP, hts.c:1833-1875 (integers, commas, e-notation
 * and fractions are accepted by the reference; the oracle keeps digits/commas/[eE]/'.'). */
static long long parse_decimal(const char* s, const char** end) {
    long long n = 0; int decimals = 0, e = 0, lost = 0; char sign = '+';
    while (*s == ' ' || *s == '\t') ++s;
    if (*s == '+' || *s == '-') sign = *s++;
    while (*s) { if (*s >= '0' && *s <= '9') n = 10 * n + (*s++ - '0'); else if (*s == ',') ++s; else break; }
    if (*s == '.') { ++s; while (*s >= '0' && *s <= '9') { decimals++; n = 10 * n + (*s++ - '0'); } }
    if (*s == 'E' || *s == 'e') { char* t; e = (int)strtol(s + 1, &t, 10); s = t; }
    e -= decimals;
    while (e > 0) { n *= 10; e--; }
    while (e < 0) { lost += n % 10; n /= 10; e++; }
    (void)lost;
    if (end) *end = s;
    return sign == '+' ? n : -n;
}

static int name2id(const jxo_t* o, const char* name, size_t len) {
    int id = -1;   /* duplicates: last index wins (sam.c:262-277) */
    for (int32_t i = 0; i < o->n_contig; ++i)
        if (strlen(o->contig[i]) == len && !memcmp(o->contig[i], name, len)) id = i;
    return id;
}

static int cmp_pair(const void* a, const void* b) {
    uint64_t x = ((const pair64*)a)->u, y = ((const pair64*)b)->u;
    return x < y ? -1 : x > y;
}

/* identify_junctions_from_BAM, junctions_extractor.cc:500-535. */
int jxo_extract_bam(jxo_t* o, const char* bam, const char* region, const char** err) {
    static const char* E_OPEN = "Unable to open BAM/SAM file.\n\n";
    static const char* E_IDX = "Unable to open BAM/SAM index. Make sure alignments are indexed\n\n";
    static const char* E_ITER = "Unable to iterate to region within BAM.\n\n";
    if (err) *err = NULL;
    int fd = open(bam, O_RDONLY);
    if (fd < 0) { if (err) *err = E_OPEN; return 1; }                       /* :503-506 */
    struct stat st; fstat(fd, &st);
    bgzf_t* fp = (bgzf_t*)calloc(1, sizeof *fp);
    fp->size = (size_t)st.st_size;
    fp->file = fp->size ? (const uint8_t*)mmap(NULL, fp->size, PROT_READ, MAP_PRIVATE, fd, 0) : NULL;
    close(fd);
    int rc = 1; bai_t* idx = NULL; rec_t rec; memset(&rec, 0, sizeof rec);
    pair64* off = NULL;
    if (fp->size && fp->file == MAP_FAILED) { if (err) *err = E_OPEN; free(fp); return 1; }
    idx = bai_find(bam);                                                    /* :508-512 */
    if (!idx) { if (err) *err = E_IDX; goto done; }
    /* sam_hdr_read :514, sam.c:114-223 */
    {
        char magic[4]; int32_t l_text, n_ref;
        bgzf_seek_(fp, 0);
        if (bgzf_read_(fp, magic, 4) != 4 || memcmp(magic, "BAM\1", 4)) { if (err) *err = E_ITER; goto done; }
        if (bgzf_read_(fp, &l_text, 4) != 4) { if (err) *err = E_ITER; goto done; }
        char* text = (char*)malloc((size_t)l_text + 1);
        bgzf_read_(fp, text, (size_t)l_text); free(text);
        if (bgzf_read_(fp, &n_ref, 4) != 4) { if (err) *err = E_ITER; goto done; }
        char** names = (char**)calloc(n_ref > 0 ? n_ref : 1, sizeof(char*));
        for (int32_t i = 0; i < n_ref; ++i) {
            int32_t l_name, l_ref;
            bgzf_read_(fp, &l_name, 4);
            names[i] = (char*)calloc((size_t)l_name + 1, 1);
            bgzf_read_(fp, names[i], (size_t)l_name);
            bgzf_read_(fp, &l_ref, 4);
        }
        jxo_set_contigs(o, n_ref, (const char* const*)names);
        for (int32_t i = 0; i < n_ref; ++i) free(names[i]);
        free(names);
    }
    if (!strcmp(region, ".")) {
        /* HTS_IDX_START, hts.c:1721-1731; read_rest path hts.c:1928-1939 */
        uint64_t off0 = (uint64_t)-1;
        for (int32_t i = 0; i < idx->n_ref; ++i) {
            bin_t* m = find_bin(&idx->ref[i], META_BIN);
            if (m && m->n > 0 && off0 > m->list[0].u) off0 = m->list[0].u;
        }
        if (off0 == (uint64_t)-1 && idx->n_no_coor) off0 = 0;
        if (off0 == (uint64_t)-1) { if (err) *err = E_ITER; goto done; }
        if (off0) bgzf_seek_(fp, off0);
        while (read_rec(fp, &rec) >= 0) feed(o, &rec);
        rc = 0; goto done;
    }
    if (!strcmp(region, "*")) {
        /* HTS_IDX_NOCOOR, hts.c:1733-1741: unmapped tail; reads there can still carry a CIGAR */
        uint64_t off0 = (uint64_t)-1;
        if (idx->n_ref > 0) {
            bin_t* m = find_bin(&idx->ref[idx->n_ref - 1], META_BIN);
            if (m && m->n > 0) off0 = m->list[0].v;
        }
        if (off0 == (uint64_t)-1 && idx->n_no_coor) off0 = 0;
        if (off0 == (uint64_t)-1) { if (err) *err = E_ITER; goto done; }
        if (off0) bgzf_seek_(fp, off0);
        while (read_rec(fp, &rec) >= 0) feed(o, &rec);
        rc = 0; goto done;
    }
    {
        /* hts_parse_reg hts.c:1877-1895 + hts_itr_querys :1897-1922 */
        int tid; long long beg, end;
        const char* colon = strrchr(region, ':');
        int parsed = 1;
        if (!colon) { beg = 0; end = INT_MAX; colon = region + strlen(region); }
        else {
            const char* hy;
            beg = parse_decimal(colon + 1, &hy) - 1;
            if (beg < 0) beg = 0;
            if (*hy == '\0') end = INT_MAX;
            else if (*hy == '-') end = parse_decimal(hy + 1, NULL);
            else parsed = 0;
            if (parsed && beg >= end) parsed = 0;
        }
        beg = (int)beg; end = (int)end;
        if (parsed) tid = name2id(o, region, (size_t)(colon - region));
        else { tid = name2id(o, region, strlen(region)); beg = 0; end = INT_MAX; }
        if (tid < 0) { if (err) *err = E_ITER; goto done; }
        /* hts_itr_query hts.c:1749-1808 */
        if (end < beg) { if (err) *err = E_ITER; goto done; }
        if (tid >= idx->n_ref) { if (err) *err = E_ITER; goto done; }
        refidx_t* r = &idx->ref[tid];
        /* NB the reference tests `bidx == NULL`; BAI-loaded indexes always have a (maybe empty) map */
        uint64_t min_off = 0;
        {
            uint32_t bin = 4681u + (uint32_t)(beg >> 14); bin_t* k = NULL;
            do {
                k = find_bin(r, bin);
                if (k) break;
                uint32_t parent = (bin - 1) >> 3, first = (parent << 3) + 1;
                if (bin > first) --bin; else bin = parent;
            } while (bin);
            if (bin == 0) k = find_bin(r, 0);
            min_off = k ? k->loff : 0;
        }
        /* reg2bins hts.c:1690-1706 */
        size_t n_off = 0, m_off = 0;
        {
            long long b = beg, e = end; int s = 14 + 15;
            if (b < e) {
                if (e >= 1ll << s) e = 1ll << s;
                --e;
                int t = 0;
                for (int l = 0; l <= 5; s -= 3, t += 1 << (3 * l), ++l) {
                    for (long long i = t + (b >> s); i <= t + (e >> s); ++i) {
                        bin_t* k = find_bin(r, (uint32_t)i);
                        if (!k) continue;
                        for (int32_t j = 0; j < k->n; ++j) if (k->list[j].v > min_off) {
                            if (n_off == m_off) { m_off = m_off ? m_off * 2 : 16; off = (pair64*)realloc(off, m_off * sizeof(pair64)); }
                            off[n_off++] = k->list[j];
                        }
                    }
                }
            }
        }
        if (n_off == 0) { rc = 0; goto done; }       /* iterator with no chunks: zero reads */
        qsort(off, n_off, sizeof(pair64), cmp_pair);
        size_t l = 0;
        for (size_t i = 1; i < n_off; ++i) if (off[l].v < off[i].v) off[++l] = off[i];
        n_off = l + 1;
        for (size_t i = 1; i < n_off; ++i) if (off[i - 1].v >= off[i].u) off[i - 1].v = off[i].u;
        l = 0;
        for (size_t i = 1; i < n_off; ++i) {
            if (off[l].v >> 16 == off[i].u >> 16) off[l].v = off[i].v; else off[++l] = off[i];
        }
        n_off = l + 1;
        /* hts_itr_next region path hts.c:1941-1963 */
        long long ci = -1; uint64_t curr_off = 0;
        for (;;) {
            if (curr_off == 0 || curr_off >= off[ci].v) {
                if (ci == (long long)n_off - 1) break;
                if (ci < 0 || off[ci].v != off[ci + 1].u) { bgzf_seek_(fp, off[ci + 1].u); curr_off = bgzf_tell_(fp); }
                ++ci;
            }
            if (read_rec(fp, &rec) < 0) break;
            curr_off = bgzf_tell_(fp);
            int32_t rend = rec_endpos(&rec);
            if (rec.tid != tid || rec.pos >= end) break;
            if (rend > beg && end > rec.pos) feed(o, &rec);
        }
        rc = 0;
    }
done:
    free(off); free(rec.data); bai_free(idx);
    if (fp->file && fp->size) munmap((void*)fp->file, fp->size);
    free(fp);
    if (rc == 0 && o->fasta_error) { if (err) *err = o->errbuf; return 1; }   /* get_reference_sequence threw (:553-555) */
    return rc;
}
