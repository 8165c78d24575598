/*
 * omc_blob.h -- tiny named-array container used to move a fully initialised ompMC problem
 * (physics tables, geometry, regions, source) from the container that has /root/reference and its
 * data files to machines that do not (the GPU box).  TEST INFRASTRUCTURE (oracle/).
 *
 * Layout (little endian):  "OMCBLOB1" | u32 narrays | narrays x { char name[32]; u32 dtype; u32 pad;
 * u64 count; payload padded to 8 bytes }.   dtype 0 = float64, 1 = int32.
 * The python reader is ompmc_b200/problem.py:load_blob().
 */
#ifndef OMC_BLOB_H
#define OMC_BLOB_H
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define OMC_BLOB_F64 0u
#define OMC_BLOB_I32 1u
#define OMC_BLOB_MAXARR 256

typedef struct omc_blob_entry { char name[32]; uint32_t dtype; uint64_t count; void *data; } omc_blob_entry;
typedef struct omc_blob { int n; omc_blob_entry e[OMC_BLOB_MAXARR]; } omc_blob;

static inline void omc_blob_add(omc_blob *b, const char *name, uint32_t dtype, uint64_t count, const void *data) {
    if (b->n >= OMC_BLOB_MAXARR) { fprintf(stderr, "omc_blob: too many arrays\n"); exit(1); }
    omc_blob_entry *e = &b->e[b->n++];
    memset(e->name, 0, sizeof e->name);
    strncpy(e->name, name, sizeof e->name - 1);
    e->dtype = dtype; e->count = count;
    size_t sz = (size_t)count * (dtype == OMC_BLOB_F64 ? 8 : 4);
    e->data = malloc(sz ? sz : 1);
    memcpy(e->data, data, sz);
}
static inline void omc_blob_add_f64(omc_blob *b, const char *name, uint64_t n, const double *d) { omc_blob_add(b, name, OMC_BLOB_F64, n, d); }
static inline void omc_blob_add_i32(omc_blob *b, const char *name, uint64_t n, const int *d) { omc_blob_add(b, name, OMC_BLOB_I32, n, d); }
static inline void omc_blob_add_scalar(omc_blob *b, const char *name, double v) { omc_blob_add(b, name, OMC_BLOB_F64, 1, &v); }
static inline void omc_blob_add_iscalar(omc_blob *b, const char *name, int v) { omc_blob_add(b, name, OMC_BLOB_I32, 1, &v); }

static inline int omc_blob_write(const omc_blob *b, const char *path) {
    FILE *fp = fopen(path, "wb");
    if (!fp) return -1;
    uint32_t n = (uint32_t)b->n, pad = 0;
    fwrite("OMCBLOB1", 1, 8, fp); fwrite(&n, 4, 1, fp);
    for (int i = 0; i < b->n; i++) {
        const omc_blob_entry *e = &b->e[i];
        size_t sz = (size_t)e->count * (e->dtype == OMC_BLOB_F64 ? 8 : 4);
        fwrite(e->name, 1, 32, fp); fwrite(&e->dtype, 4, 1, fp); fwrite(&pad, 4, 1, fp);
        fwrite(&e->count, 8, 1, fp); fwrite(e->data, 1, sz, fp);
        if (sz % 8) fwrite(&pad, 1, 8 - sz % 8, fp);
    }
    fclose(fp);
    return 0;
}
static inline int omc_blob_read(omc_blob *b, const char *path) {
    FILE *fp = fopen(path, "rb");
    if (!fp) return -1;
    char magic[8]; uint32_t n = 0, pad;
    if (fread(magic, 1, 8, fp) != 8 || memcmp(magic, "OMCBLOB1", 8) || fread(&n, 4, 1, fp) != 1) { fclose(fp); return -2; }
    b->n = 0;
    for (uint32_t i = 0; i < n && i < OMC_BLOB_MAXARR; i++) {
        omc_blob_entry *e = &b->e[b->n];
        if (fread(e->name, 1, 32, fp) != 32 || fread(&e->dtype, 4, 1, fp) != 1 || fread(&pad, 4, 1, fp) != 1 ||
            fread(&e->count, 8, 1, fp) != 1) { fclose(fp); return -3; }
        size_t sz = (size_t)e->count * (e->dtype == OMC_BLOB_F64 ? 8 : 4);
        size_t psz = (sz + 7) & ~(size_t)7;
        e->data = malloc(psz ? psz : 8);
        if (fread(e->data, 1, psz, fp) != psz) { fclose(fp); return -4; }
        b->n++;
    }
    fclose(fp);
    return 0;
}
static inline const omc_blob_entry *omc_blob_find(const omc_blob *b, const char *name) {
    for (int i = 0; i < b->n; i++) if (!strncmp(b->e[i].name, name, 32)) return &b->e[i];
    fprintf(stderr, "omc_blob: array '%s' not found\n", name);
    exit(1);
}
static inline const double *omc_blob_f64(const omc_blob *b, const char *name) { return (const double *)omc_blob_find(b, name)->data; }
static inline const int *omc_blob_i32(const omc_blob *b, const char *name) { return (const int *)omc_blob_find(b, name)->data; }
static inline void omc_blob_free(omc_blob *b) { for (int i = 0; i < b->n; i++) free(b->e[i].data); b->n = 0; }
#endif
