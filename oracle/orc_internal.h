/* orc_internal.h — TEST INFRASTRUCTURE ONLY (see gofr_oracle.h). Shared declarations of the CPU oracle. */
#ifndef ORC_INTERNAL_H
#define ORC_INTERNAL_H
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* growable byte buffer (models bytes.Buffer / bufio) */
typedef struct {
    uint8_t* p;
    size_t n, cap;
} obuf;

static inline void ob_init(obuf* b) { b->p = NULL; b->n = b->cap = 0; }
static inline void ob_free(obuf* b) { free(b->p); b->p = NULL; b->n = b->cap = 0; }
static inline void ob_reserve(obuf* b, size_t extra) {
    if (b->n + extra > b->cap) {
        size_t c = b->cap ? b->cap * 2 : 256;
        while (c < b->n + extra) c *= 2;
        b->p = (uint8_t*)realloc(b->p, c);
        b->cap = c;
    }
}
static inline void ob_put(obuf* b, const void* s, size_t n) {
    ob_reserve(b, n);
    if (n) memcpy(b->p + b->n, s, n);
    b->n += n;
}
static inline void ob_putc(obuf* b, uint8_t c) { ob_reserve(b, 1); b->p[b->n++] = c; }
static inline void ob_puts(obuf* b, const char* s) { ob_put(b, s, strlen(s)); }

/* field kinds (numeric values match include/gofr_b200.h) */
enum { F_INT64 = 1, F_INT32 = 2, F_BOOL = 3, F_STRING = 4, F_INT = 5, F_FLOAT64 = 6, F_STRUCT = 7, F_UINT64 = 8, F_BYTES = 9, F_FLOAT32 = 10, F_TIME = 11 };
enum { C_VALUE = 0, C_PTR = 1, C_SLICE = 2, C_MAP = 3, C_SLICE_PTR = 4 }; /* T, *T, []T, map[string]T, []*T */
enum { FIELD_BARE = 1 };                                 /* one-field schema standing for the field's own type */

typedef struct {
    char* go_name;
    char* json_name;
    int kind;
    int omitempty;
    int container;
    int flags;
    int elem_schema; /* F_STRUCT: schema id */
} orc_field;

typedef struct {
    int id;
    char* go_type; /* reflect.Type.String(), e.g. "main.Person" */
    int n_fields;
    orc_field* f;
} orc_schema;

/* a decoded struct value */
typedef struct {
    int64_t i;        /* INT*, BOOL */
    const uint8_t* s; /* STRING */
    int sn;
    uint8_t* owned; /* non-NULL if s was allocated by bind */
} orc_value;

/* encoding/json pieces */
void orc_enc_string(obuf* b, const uint8_t* s, size_t n);
void orc_enc_int(obuf* b, int64_t v);
void orc_enc_struct(obuf* b, const orc_schema* sc, const orc_value* v);

/* json.Unmarshal into struct.  Returns 0 and fills vals (caller frees owned), or 1 and writes err.Error() to err. */
int orc_unmarshal_struct(const orc_schema* sc, const uint8_t* body, size_t n, orc_value* vals, obuf* err);

const char* orc_go_kind_name(int kind);

/* values beyond flat structs (orc_value.c) */
struct orc_table;
const orc_schema* orc_find_schema(const struct orc_table* t, int id);
int orc_float_text(double x, char* out); /* encoding/json float64 text; 0 = NaN / Inf (not encodable) */
int orc_float32_text(float x, char* out); /* the same for a float32 (shortest digits that identify the float32) */
int orc_schema_is_flat(const orc_schema* sc);
int orc_schema_fixed_words(const struct orc_table* t, const orc_schema* sc);
int orc_enc_row(obuf* b, const struct orc_table* t, const orc_schema* sc, const uint8_t* fixed, size_t fixed_avail,
                const uint8_t* var, const uint8_t* end);

#endif
