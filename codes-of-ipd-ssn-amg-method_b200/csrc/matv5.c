/* matv5.c -- reader for the MAT-file level 5 inputs the reference ships (Class1/InputData/data1-500.mat,
 * Class2/InputData/data4-500.mat: "MATLAB 5.0 MAT-file", little endian, every variable a zlib-compressed
 * miMATRIX element whose numeric data MATLAB stored in the smallest integer type that holds it).
 * the input path of a standalone (non-MATLAB, non-Python) run.  Host code, plain C,
 * built into libssnmat.so (zlib is the only dependency); declared in include/ssnamg_io.h.
 *
 * Supported: real full numeric arrays of any class (double ... uint64, stored as any miINT8 ... miDOUBLE
 * type), compressed or not, 2-D.  Everything else (sparse, char, logical-only, struct, cell, complex) is
 * listed with its name and reported as unsupported when read.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <zlib.h>

#include "../../include/ssnamg_io.h"

enum { miINT8 = 1, miUINT8 = 2, miINT16 = 3, miUINT16 = 4, miINT32 = 5, miUINT32 = 6, miSINGLE = 7, miDOUBLE = 9,
       miINT64 = 12, miUINT64 = 13, miMATRIX = 14, miCOMPRESSED = 15 };
enum { mxDOUBLE_CLASS = 6, mxUINT64_CLASS = 15 };

typedef struct {
    char name[64];
    int64_t rows, cols;
    int supported;              /* real full numeric 2-D array */
    uint32_t dtype;             /* stored type of the real part */
    const uint8_t* data;        /* points into `buf` */
    uint32_t nbytes;
    uint8_t* buf;               /* inflated element (owned) or NULL when the element was stored uncompressed */
} ssn_mat_var;

struct ssn_mat {
    uint8_t* file; size_t size;
    ssn_mat_var* vars; int count, cap;
};

static uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }

/* tag at p: returns the data pointer, type and byte count; *adv = bytes to the next element (padded to 8) */
static const uint8_t* read_tag(const uint8_t* p, const uint8_t* end, uint32_t* type, uint32_t* nbytes, size_t* adv) {
    if (end - p < 8) return NULL;
    const uint32_t w0 = rd32(p);
    if (w0 >> 16) {                                 /* small data element: type in the low, size in the high half-word */
        *type = w0 & 0xffffu; *nbytes = w0 >> 16; *adv = 8;
        return (*nbytes <= 4) ? p + 4 : NULL;
    }
    *type = w0; *nbytes = rd32(p + 4);
    if ((size_t)(end - p - 8) < *nbytes) return NULL;
    *adv = 8 + (size_t)*nbytes + ((8 - (*nbytes & 7u)) & 7u);
    return p + 8;
}

static int type_size(uint32_t t) {
    switch (t) {
        case miINT8: case miUINT8: return 1;
        case miINT16: case miUINT16: return 2;
        case miINT32: case miUINT32: case miSINGLE: return 4;
        case miDOUBLE: case miINT64: case miUINT64: return 8;
        default: return 0;
    }
}

/* parses one miMATRIX payload [p, p+n) into v (pointers into the payload) */
static void parse_matrix(const uint8_t* p, uint32_t n, ssn_mat_var* v) {
    const uint8_t* end = p + n;
    uint32_t type, nb; size_t adv;
    v->supported = 0; v->rows = v->cols = -1; v->name[0] = 0;
    const uint8_t* d = read_tag(p, end, &type, &nb, &adv);                 /* array flags */
    if (!d || type != miUINT32 || nb < 8) return;
    const uint32_t flags = rd32(d);
    const uint32_t cls = flags & 0xffu;
    const int is_complex = (flags >> 11) & 1;
    p += adv;
    d = read_tag(p, end, &type, &nb, &adv);                                /* dimensions */
    if (!d || type != miINT32 || nb < 8) return;
    const int ndim = (int)(nb / 4);
    int64_t dims[2] = {(int32_t)rd32(d), (int32_t)rd32(d + 4)};
    for (int k = 2; k < ndim; ++k) if ((int32_t)rd32(d + 4 * k) != 1) dims[0] = -1;
    p += adv;
    d = read_tag(p, end, &type, &nb, &adv);                                /* name */
    if (!d || type != miINT8) return;
    { const uint32_t len = nb < sizeof(v->name) - 1 ? nb : (uint32_t)sizeof(v->name) - 1; memcpy(v->name, d, len); v->name[len] = 0; }
    p += adv;
    v->rows = dims[0]; v->cols = dims[1];
    if (cls < mxDOUBLE_CLASS || cls > mxUINT64_CLASS || is_complex || dims[0] < 0 || dims[1] < 0) return;   /* cell, struct, char, sparse, complex, N-D */
    d = read_tag(p, end, &type, &nb, &adv);                                /* real part */
    if (!d) return;
    const int ts = type_size(type);
    if (!ts) return;
    {   /* file-supplied dimensions: bound each by the element count before multiplying (no int64 overflow) */
        const int64_t nel = (int64_t)nb / ts;
        if (dims[0] == 0 || dims[1] == 0) { if (nb != 0) return; }
        else {
            if (dims[0] > nel || dims[1] > nel / dims[0]) return;
            if ((int64_t)nb != dims[0] * dims[1] * ts) return;
        }
    }
    v->dtype = type; v->data = d; v->nbytes = nb; v->supported = 1;
}

static int push_var(ssn_mat* m, const ssn_mat_var* v) {
    if (m->count == m->cap) {
        const int nc = m->cap ? 2 * m->cap : 16;
        ssn_mat_var* nv = (ssn_mat_var*)realloc(m->vars, sizeof(ssn_mat_var) * (size_t)nc);
        if (!nv) return SSN_MAT_E_NOMEM;
        m->vars = nv; m->cap = nc;
    }
    m->vars[m->count++] = *v;
    return SSN_MAT_OK;
}

int ssn_mat_open(const char* path, ssn_mat** out) {
    if (!path || !out) return SSN_MAT_E_INVALID;
    *out = NULL;
    FILE* f = fopen(path, "rb");
    if (!f) return SSN_MAT_E_IO;
    fseek(f, 0, SEEK_END); const long sz = ftell(f); fseek(f, 0, SEEK_SET);
    if (sz < 128) { fclose(f); return SSN_MAT_E_FORMAT; }
    ssn_mat* m = (ssn_mat*)calloc(1, sizeof(ssn_mat));
    if (!m) { fclose(f); return SSN_MAT_E_NOMEM; }
    m->file = (uint8_t*)malloc((size_t)sz); m->size = (size_t)sz;
    if (!m->file || fread(m->file, 1, (size_t)sz, f) != (size_t)sz) { fclose(f); ssn_mat_close(m); return SSN_MAT_E_IO; }
    fclose(f);
    /* header: 116 bytes of text, 8 bytes subsystem offset, version 0x0100, endian indicator "IM" (little endian files) */
    if (memcmp(m->file, "MATLAB 5.0 MAT-file", 19) != 0 || m->file[126] != 'I' || m->file[127] != 'M') { ssn_mat_close(m); return SSN_MAT_E_FORMAT; }
    const uint8_t* p = m->file + 128; const uint8_t* end = m->file + m->size;
    while (p < end) {
        uint32_t type, nb; size_t adv;
        if (end - p < 8) break;
        type = rd32(p); nb = rd32(p + 4);                                   /* top-level elements use the long tag */
        if ((size_t)(end - p - 8) < nb) { ssn_mat_close(m); return SSN_MAT_E_FORMAT; }
        ssn_mat_var v; memset(&v, 0, sizeof(v));
        if (type == miCOMPRESSED) {
            adv = 8 + (size_t)nb;                                           /* compressed elements are not padded */
            /* inflate: the first 8 bytes give the tag of the inner element, hence its size */
            uint8_t head[8]; z_stream zs; memset(&zs, 0, sizeof(zs));
            if (inflateInit(&zs) != Z_OK) { ssn_mat_close(m); return SSN_MAT_E_ZLIB; }
            zs.next_in = (Bytef*)(p + 8); zs.avail_in = nb; zs.next_out = head; zs.avail_out = 8;
            int zr = inflate(&zs, Z_SYNC_FLUSH);
            if ((zr != Z_OK && zr != Z_STREAM_END) || zs.avail_out != 0) { inflateEnd(&zs); ssn_mat_close(m); return SSN_MAT_E_ZLIB; }
            const uint32_t itype = rd32(head), inb = rd32(head + 4);
            uint8_t* buf = (uint8_t*)malloc((size_t)inb + 8);
            if (!buf) { inflateEnd(&zs); ssn_mat_close(m); return SSN_MAT_E_NOMEM; }
            memcpy(buf, head, 8);
            zs.next_out = buf + 8; zs.avail_out = inb;
            zr = inflate(&zs, Z_FINISH);
            inflateEnd(&zs);
            if (zs.avail_out != 0 || (zr != Z_STREAM_END && zr != Z_OK && zr != Z_BUF_ERROR)) { free(buf); ssn_mat_close(m); return SSN_MAT_E_ZLIB; }   /* the element must inflate to exactly the size its tag states */
            v.buf = buf;
            if (itype == miMATRIX) parse_matrix(buf + 8, inb, &v);
            else { free(buf); p += adv; continue; }
        } else if (type == miMATRIX) {
            adv = 8 + (size_t)nb + ((8 - (nb & 7u)) & 7u);
            parse_matrix(p + 8, nb, &v);
        } else {
            adv = 8 + (size_t)nb + ((8 - (nb & 7u)) & 7u);
            p += adv; continue;
        }
        const int st = push_var(m, &v);
        if (st != SSN_MAT_OK) { free(v.buf); ssn_mat_close(m); return st; }
        p += adv;
    }
    *out = m;
    return SSN_MAT_OK;
}

void ssn_mat_close(ssn_mat* m) {
    if (!m) return;
    for (int i = 0; i < m->count; ++i) free(m->vars[i].buf);
    free(m->vars); free(m->file); free(m);
}

int ssn_mat_count(const ssn_mat* m) { return m ? m->count : 0; }

const char* ssn_mat_name(const ssn_mat* m, int i) { return (m && i >= 0 && i < m->count) ? m->vars[i].name : NULL; }

int ssn_mat_find(const ssn_mat* m, const char* name) {
    if (!m || !name) return -1;
    for (int i = 0; i < m->count; ++i) if (strcmp(m->vars[i].name, name) == 0) return i;
    return -1;
}

int ssn_mat_dims(const ssn_mat* m, int i, int64_t* rows, int64_t* cols) {
    if (!m || i < 0 || i >= m->count) return SSN_MAT_E_INVALID;
    if (rows) *rows = m->vars[i].rows;
    if (cols) *cols = m->vars[i].cols;
    return m->vars[i].supported ? SSN_MAT_OK : SSN_MAT_E_UNSUPPORTED;
}

int ssn_mat_read_double(const ssn_mat* m, int i, double* out) {
    if (!m || !out || i < 0 || i >= m->count) return SSN_MAT_E_INVALID;
    const ssn_mat_var* v = &m->vars[i];
    if (!v->supported) return SSN_MAT_E_UNSUPPORTED;
    const int64_t n = v->rows * v->cols;
    const uint8_t* d = v->data;
    for (int64_t k = 0; k < n; ++k) {
        switch (v->dtype) {
            case miINT8:   out[k] = (double)((const int8_t*)d)[k]; break;
            case miUINT8:  out[k] = (double)d[k]; break;
            case miINT16:  { int16_t t; memcpy(&t, d + 2 * k, 2); out[k] = (double)t; } break;
            case miUINT16: { uint16_t t; memcpy(&t, d + 2 * k, 2); out[k] = (double)t; } break;
            case miINT32:  { int32_t t; memcpy(&t, d + 4 * k, 4); out[k] = (double)t; } break;
            case miUINT32: { uint32_t t; memcpy(&t, d + 4 * k, 4); out[k] = (double)t; } break;
            case miSINGLE: { float t; memcpy(&t, d + 4 * k, 4); out[k] = (double)t; } break;
            case miDOUBLE: { double t; memcpy(&t, d + 8 * k, 8); out[k] = t; } break;
            case miINT64:  { int64_t t; memcpy(&t, d + 8 * k, 8); out[k] = (double)t; } break;
            case miUINT64: { uint64_t t; memcpy(&t, d + 8 * k, 8); out[k] = (double)t; } break;
            default: return SSN_MAT_E_UNSUPPORTED;
        }
    }
    return SSN_MAT_OK;
}

const char* ssn_mat_strerror(int status) {
    switch (status) {
        case SSN_MAT_OK:            return "ok";
        case SSN_MAT_E_IO:          return "cannot read the file";
        case SSN_MAT_E_FORMAT:      return "not a little-endian level 5 MAT-file (or truncated)";
        case SSN_MAT_E_ZLIB:        return "a compressed element does not inflate";
        case SSN_MAT_E_NOMEM:       return "out of memory";
        case SSN_MAT_E_INVALID:     return "invalid argument, missing variable or inconsistent sizes";
        case SSN_MAT_E_UNSUPPORTED: return "variable is not a real full numeric 2-D array";
        default:                    return "unknown status";
    }
}

/* numel-checked copy of variable `name` into a fresh buffer; *out = NULL when absent and optional */
static int load_vec(const ssn_mat* m, const char* name, int64_t want, int optional, double** out) {
    *out = NULL;
    const int i = ssn_mat_find(m, name);
    if (i < 0) return optional ? SSN_MAT_OK : SSN_MAT_E_INVALID;
    int64_t r, c;
    const int st = ssn_mat_dims(m, i, &r, &c);
    if (st != SSN_MAT_OK) return st;
    if (want >= 0 && r * c != want) return SSN_MAT_E_INVALID;
    double* buf = (double*)malloc(sizeof(double) * (size_t)(r * c > 0 ? r * c : 1));
    if (!buf) return SSN_MAT_E_NOMEM;
    const int st2 = ssn_mat_read_double(m, i, buf);
    if (st2 != SSN_MAT_OK) { free(buf); return st2; }
    *out = buf;
    return SSN_MAT_OK;
}

static int64_t numel_of(const ssn_mat* m, const char* name) {
    int64_t r, c;
    const int i = ssn_mat_find(m, name);
    if (i < 0 || ssn_mat_dims(m, i, &r, &c) != SSN_MAT_OK) return -1;
    return r * c;
}

static int load_scalar(const ssn_mat* m, const char* name, double* out) {
    double* b; const int st = load_vec(m, name, 1, 0, &b);
    if (st != SSN_MAT_OK) return st;
    *out = b[0]; free(b);
    return SSN_MAT_OK;
}

int ssn_problem_load(const char* path, ssn_problem* pb) {
    if (!pb) return SSN_MAT_E_INVALID;
    memset(pb, 0, sizeof(*pb));
    pb->mu = (double)NAN;
    ssn_mat* m = NULL;
    int st = ssn_mat_open(path, &m);
    if (st != SSN_MAT_OK) return st;
    double v;
    pb->m = (load_scalar(m, "m", &v) == SSN_MAT_OK) ? (int64_t)v : numel_of(m, "l");
    pb->n = (load_scalar(m, "n", &v) == SSN_MAT_OK) ? (int64_t)v : numel_of(m, "r");
    if (pb->m <= 0 || pb->n <= 0) { ssn_mat_close(m); return SSN_MAT_E_INVALID; }
    const int64_t mn = pb->m * pb->n;
    if ((st = load_vec(m, "c", mn, 0, &pb->c)) != SSN_MAT_OK) goto fail;
    if ((st = load_vec(m, "r", pb->n, 0, &pb->r)) != SSN_MAT_OK) goto fail;
    if ((st = load_vec(m, "l", pb->m, 0, &pb->l)) != SSN_MAT_OK) goto fail;
    if ((st = load_vec(m, "p", pb->m, 0, &pb->p)) != SSN_MAT_OK) goto fail;
    if ((st = load_vec(m, "q", pb->n, 0, &pb->q)) != SSN_MAT_OK) goto fail;
    if ((st = load_vec(m, "gama", mn, 1, &pb->gama)) != SSN_MAT_OK) goto fail;
    if ((st = load_vec(m, "phi", mn, 1, &pb->phi)) != SSN_MAT_OK) goto fail;
    if (ssn_mat_find(m, "mu") >= 0 && (st = load_scalar(m, "mu", &pb->mu)) != SSN_MAT_OK) goto fail;
    ssn_mat_close(m);
    return SSN_MAT_OK;
fail:
    ssn_mat_close(m);
    ssn_problem_free(pb);
    return st;
}

void ssn_problem_free(ssn_problem* pb) {
    if (!pb) return;
    free(pb->c); free(pb->r); free(pb->l); free(pb->p); free(pb->q); free(pb->gama); free(pb->phi);
    memset(pb, 0, sizeof(*pb));
}
