// proto_nested_decode_device.cuh — proto3 decoder for message types with nested and repeated fields
// (gofr_proto_decode_nested_device): dec(in) of a unary handler (examples/grpc-server/grpc/hello_grpc.pb.go:73-89) — the
// 5-byte length-prefixed-message header checked, proto.Unmarshal into the message type, the values out as a row in the layout
// gofr_proto_encode_nested_device reads (include/gofr_b200.h "Row format").  protobuf-go v1.32.0 rules restated
// (impl/decode.go, codec_field.go, codec_gen.go), on top of the flat decoder's (grpc_device.cuh):
//   * a known field arriving with another wire type than its own is an unknown field — except that a repeated numeric
//     scalar is accepted BOTH packed (wire type 2, elements back to back) and unpacked (one element per tag), mixed freely,
//     elements in wire order;
//   * singular scalars / strings: the last occurrence wins; every occurrence of a string must be valid UTF-8;
//   * a singular message field occurring more than once is MERGED by protobuf-go; that is not done here: the frame comes
//     back with GOFR_GRPC_DEFER ("valid, but decode it on the host") — canonical encoders write the field once;
//   * repeated strings / bytes / messages: one element per occurrence, in wire order; unknown fields and balanced groups
//     are skipped at every level; nested messages are only parsed where the schema knows them.
// Three walks per frame, none recursive (explicit stacks, at most kPbnMaxDepth message levels):
//   validate   wire order: framing of every level that the schema knows, UTF-8, packed payloads, repeated singular messages
//   size       field order: the row's byte count
//   emit       field order again, this time storing — the fixed part of a message is written where its parent put it, the
//              variable part is appended; byte stores, no Writer: a row is not produced front to back.
// Status precedence when a frame has several problems: malformed > invalid UTF-8 > DEFER (protobuf-go reports whichever
// comes first in wire order; either way the RPC fails).  __host__ __device__: tests/emu runs it on the CPU.
#pragma once
#include "proto_nested_device.cuh"

namespace gofr {

struct PdnRow {
    uint32_t status;   // GOFR_GRPC_*
    uint32_t out_len;  // bytes of the row (multiple of 4; 0 unless GOFR_GRPC_OK)
};

// the scalar's row words from its wire value (the flat decoder's rules)
GOFR_HD void pdn_scalar_words(uint32_t cls, uint32_t wt, uint64_t v, uint32_t* w0, uint32_t* w1) {
    *w1 = 0;
    if (wt != 0 || ((cls & PC_64) && !(cls & PC_ZIGZAG))) { *w0 = (uint32_t)v; *w1 = (uint32_t)(v >> 32); }
    else if (cls & PC_ZIGZAG) {
        if (cls & PC_64) { const uint64_t z = (v >> 1) ^ (uint64_t)-(int64_t)(v & 1); *w0 = (uint32_t)z; *w1 = (uint32_t)(z >> 32); }
        else { const uint32_t x = (uint32_t)v; *w0 = (x >> 1) ^ (uint32_t)-(int32_t)(x & 1); }
    } else if (cls & PC_BOOL) *w0 = v != 0 ? 1u : 0u;
    else *w0 = (uint32_t)v;
    if (!(cls & PC_64)) *w1 = 0;
}

// One field of a message: tag and payload.  Returns false on malformed bytes.  wt 3 / 4 (groups) return with plen = 0.
struct PdnItem { uint32_t num, wt, poff, plen; uint64_t v; };
GOFR_HD bool pdn_next(const uint8_t* f, uint32_t* ip, uint32_t hi, PdnItem* it) {
    uint32_t i = *ip;
    uint64_t tag;
    int k = grpc_varint(f + i, hi - i, &tag);
    if (k < 0) return false;
    i += (uint32_t)k;
    const uint64_t num = tag >> 3;
    if (num == 0 || num > 0x1FFFFFFFull) return false;
    it->num = (uint32_t)num; it->wt = (uint32_t)(tag & 7); it->poff = 0; it->plen = 0; it->v = 0;
    uint64_t v = 0;
    if (it->wt == 0) {
        k = grpc_varint(f + i, hi - i, &v);
        if (k < 0) return false;
        i += (uint32_t)k;
    } else if (it->wt == 1) {
        if (hi - i < 8) return false;
        for (int q = 7; q >= 0; q--) v = v << 8 | f[i + (uint32_t)q];
        i += 8;
    } else if (it->wt == 5) {
        if (hi - i < 4) return false;
        for (int q = 3; q >= 0; q--) v = v << 8 | f[i + (uint32_t)q];
        i += 4;
    } else if (it->wt == 2) {
        k = grpc_varint(f + i, hi - i, &v);
        if (k < 0) return false;
        i += (uint32_t)k;
        if (v > hi - i) return false;
        it->poff = i;
        it->plen = (uint32_t)v;
        i += it->plen;
    } else if (it->wt != 3 && it->wt != 4) return false;
    it->v = v;
    *ip = i;
    return true;
}

// the field of message type mt that a (number, wire type) addresses, or -1: same number, and the field's own wire type or —
// for a repeated numeric scalar — either its element wire type or 2 (packed)
GOFR_HD int pdn_field_of(const PbnDesc& D, uint32_t mt, uint32_t num, uint32_t wt) {
    const PbnField* F = D.f + D.first[mt];
    for (uint32_t q = 0; q < D.count[mt]; q++) {
        if ((F[q].tag >> 3) != num) continue;
        const uint32_t own = F[q].msg != 0xFF ? 2u : (uint32_t)(F[q].cls & PC_WIRE);
        if (wt == own) return (int)q;
        if (F[q].repeated && F[q].msg == 0xFF && own != 2u && wt == 2u) return (int)q;  // packed
        return -1;
    }
    return -1;
}

// elements of a packed payload [lo, hi) of class cls: count them, or fail when the payload is not a whole number of elements
GOFR_HD bool pdn_packed_count(const uint8_t* f, uint32_t lo, uint32_t hi, uint32_t cls, uint32_t* n) {
    const uint32_t wire = cls & PC_WIRE;
    if (wire == 1) { if ((hi - lo) & 7u) return false; *n = (hi - lo) >> 3; return true; }
    if (wire == 5) { if ((hi - lo) & 3u) return false; *n = (hi - lo) >> 2; return true; }
    uint32_t c = 0;
    while (lo < hi) {
        uint64_t v;
        const int k = grpc_varint(f + lo, hi - lo, &v);
        if (k < 0) return false;
        lo += (uint32_t)k;
        c++;
    }
    *n = c;
    return true;
}

// ---- walk 1: validate, in wire order ----
GOFR_HD_NOINLINE uint32_t pdn_validate(const PbnDesc& D, const uint8_t* f, uint32_t lo, uint32_t hi) {
    struct Lv { uint32_t end; uint64_t seen; uint8_t mt; uint8_t gbase; } lv[kPbnMaxDepth + 1];
    uint32_t groups[kMaxGroupDepth];
    int depth = 0;
    uint32_t gdepth = 0, i = lo;
    bool bad_utf8 = false, defer = false;
    lv[depth].end = hi; lv[depth].seen = 0; lv[depth].mt = (uint8_t)D.root; lv[depth].gbase = 0; depth++;
    for (;;) {
        Lv& L = lv[depth - 1];
        if (i == L.end) {
            if (gdepth != L.gbase) return GOFR_GRPC_BAD_PROTO;  // a group left open inside this message
            if (--depth == 0) break;
            continue;
        }
        PdnItem it;
        if (!pdn_next(f, &i, L.end, &it)) return GOFR_GRPC_BAD_PROTO;
        if (it.wt == 3) {
            if (gdepth == kMaxGroupDepth) return GOFR_GRPC_BAD_PROTO;
            groups[gdepth++] = it.num;
            continue;
        }
        if (it.wt == 4) {
            if (gdepth == L.gbase || groups[gdepth - 1] != it.num) return GOFR_GRPC_BAD_PROTO;
            gdepth--;
            continue;
        }
        if (gdepth != L.gbase) continue;  // inside an unknown group: nothing is known
        const int q = pdn_field_of(D, L.mt, it.num, it.wt);
        if (q < 0) continue;
        const PbnField& F = D.f[D.first[L.mt] + (uint32_t)q];
        if (F.msg != 0xFF) {
            if (!F.repeated) {
                if (L.seen >> q & 1u) defer = true;
                L.seen |= 1ull << q;
            }
            if (depth > kPbnMaxDepth) return GOFR_GRPC_BAD_PROTO;  // cannot happen: the schema's depth was checked
            i = it.poff;  // descend: the payload is parsed as a message of type F.msg
            lv[depth].end = it.poff + it.plen; lv[depth].seen = 0; lv[depth].mt = F.msg; lv[depth].gbase = (uint8_t)gdepth; depth++;
            continue;
        }
        if (it.wt == 2) {
            if ((F.cls & PC_WIRE) == 2) { if ((F.cls & PC_UTF8) && !proto_utf8_ok(f + it.poff, it.plen)) bad_utf8 = true; }
            else { uint32_t n; if (!pdn_packed_count(f, it.poff, it.poff + it.plen, F.cls, &n)) return GOFR_GRPC_BAD_PROTO; }
        }
    }
    return bad_utf8 ? GOFR_GRPC_BAD_UTF8 : defer ? GOFR_GRPC_DEFER : GOFR_GRPC_OK;
}

GOFR_HD void pdn_st32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }

// ---- walks 2 and 3: field order; EMIT stores into row[] ----
// Returns the row's byte count before padding.  The frame has been validated: no check can fail here.
template <bool EMIT>
GOFR_HD_NOINLINE uint32_t pdn_walk(const PbnDesc& D, const uint8_t* f, uint32_t lo, uint32_t hi, uint8_t* row) {
    struct Fr { uint32_t lo, hi, fx, scan; uint8_t mt, fi, in_rep; } st[kPbnMaxDepth + 1];
    int depth = 0;
    uint32_t vc = (uint32_t)D.fixed_words[D.root] * 4u;  // append position of the variable part
    st[0].lo = lo; st[0].hi = hi; st[0].fx = 0; st[0].scan = 0; st[0].mt = (uint8_t)D.root; st[0].fi = 0; st[0].in_rep = 0;
    depth = 1;
    // the next top-level item of [i, hi) that addresses field q of message type mt (unknown groups skipped whole)
    auto find = [&](uint32_t mt, uint32_t q, uint32_t* ip, uint32_t end, PdnItem* out) -> bool {
        uint32_t i = *ip, g = 0;
        while (i < end) {
            PdnItem it;
            pdn_next(f, &i, end, &it);
            if (it.wt == 3) { g++; continue; }
            if (it.wt == 4) { g--; continue; }
            if (g) continue;
            if (pdn_field_of(D, mt, it.num, it.wt) == (int)q) { *out = it; *ip = i; return true; }
        }
        *ip = i;
        return false;
    };
    while (depth > 0) {
        Fr& fr = st[depth - 1];
        const PbnField* fields = D.f + D.first[fr.mt];
        if (fr.in_rep) {  // repeated message field fr.fi - 1: the next occurrence becomes the next element
            const PbnField& F = fields[fr.fi - 1];
            PdnItem it;
            if (find(fr.mt, fr.fi - 1u, &fr.scan, fr.hi, &it)) {
                const uint32_t efx = vc;
                vc += (uint32_t)D.fixed_words[F.msg] * 4u;
                Fr& ch = st[depth++];
                ch.lo = it.poff; ch.hi = it.poff + it.plen; ch.fx = efx; ch.scan = 0; ch.mt = F.msg; ch.fi = 0; ch.in_rep = 0;
            } else fr.in_rep = 0;
            continue;
        }
        if (fr.fi >= D.count[fr.mt]) { depth--; continue; }
        const uint32_t q = fr.fi;
        const PbnField& F = fields[q];
        uint32_t wo = 0;
        for (uint32_t k = 0; k < q; k++) wo += fields[k].fixed_words;
        const uint32_t fxo = fr.fx + wo * 4u;  // where this field's fixed words go
        fr.fi++;
        const uint32_t cls = F.cls;
        PdnItem it;
        uint32_t i = fr.lo;
        if (!F.repeated) {
            if (F.msg != 0xFF) {  // singular message: presence word + its fixed part inline (zeros when absent)
                const bool present = find(fr.mt, q, &i, fr.hi, &it);
                if (EMIT) {
                    pdn_st32(row + fxo, present ? 1u : 0u);
                    if (!present) for (uint32_t k = 0; k < (uint32_t)D.fixed_words[F.msg] * 4u; k++) row[fxo + 4u + k] = 0;
                }
                if (present) {
                    Fr& ch = st[depth++];
                    ch.lo = it.poff; ch.hi = it.poff + it.plen; ch.fx = fxo + 4u; ch.scan = 0; ch.mt = F.msg; ch.fi = 0; ch.in_rep = 0;
                }
            } else if ((cls & PC_WIRE) == 2) {  // string / bytes: the last occurrence
                uint32_t poff = 0, plen = 0;
                while (find(fr.mt, q, &i, fr.hi, &it)) { poff = it.poff; plen = it.plen; }
                if (EMIT) {
                    pdn_st32(row + fxo, plen);
                    for (uint32_t k = 0; k < plen; k++) row[vc + k] = f[poff + k];
                }
                vc += plen;
            } else {
                uint32_t w0 = 0, w1 = 0;
                while (find(fr.mt, q, &i, fr.hi, &it)) pdn_scalar_words(cls, it.wt, it.v, &w0, &w1);
                if (EMIT) { pdn_st32(row + fxo, w0); if (cls & PC_64) pdn_st32(row + fxo + 4u, w1); }
            }
            continue;
        }
        // repeated: the count word, the elements appended
        if (F.msg != 0xFF) {
            uint32_t n = 0;
            while (find(fr.mt, q, &i, fr.hi, &it)) n++;
            if (EMIT) pdn_st32(row + fxo, n);
            if (n) { fr.in_rep = 1; fr.scan = fr.lo; }
        } else if ((cls & PC_WIRE) == 2) {
            uint32_t n = 0;
            while (find(fr.mt, q, &i, fr.hi, &it)) {
                if (EMIT) {
                    pdn_st32(row + vc, it.plen);
                    for (uint32_t k = 0; k < it.plen; k++) row[vc + 4u + k] = f[it.poff + k];
                }
                vc += 4u + it.plen;
                n++;
            }
            if (EMIT) pdn_st32(row + fxo, n);
        } else {
            const uint32_t eb = (cls & PC_64) ? 8u : 4u, ewt = cls & PC_WIRE;
            uint32_t n = 0;
            while (find(fr.mt, q, &i, fr.hi, &it)) {
                if (it.wt == 2) {  // packed
                    uint32_t j = it.poff;
                    const uint32_t end = it.poff + it.plen;
                    while (j < end) {
                        uint64_t v = 0;
                        if (ewt == 0) { j += (uint32_t)grpc_varint(f + j, end - j, &v); }
                        else { for (int b = (int)eb - 1; b >= 0; b--) v = v << 8 | f[j + (uint32_t)b]; j += eb; }
                        if (EMIT) {
                            uint32_t w0, w1;
                            pdn_scalar_words(cls, ewt, v, &w0, &w1);
                            pdn_st32(row + vc, w0);
                            if (eb == 8) pdn_st32(row + vc + 4u, w1);
                        }
                        vc += eb;
                        n++;
                    }
                } else {
                    if (EMIT) {
                        uint32_t w0, w1;
                        pdn_scalar_words(cls, it.wt, it.v, &w0, &w1);
                        pdn_st32(row + vc, w0);
                        if (eb == 8) pdn_st32(row + vc + 4u, w1);
                    }
                    vc += eb;
                    n++;
                }
            }
            if (EMIT) pdn_st32(row + fxo, n);
        }
    }
    return vc;
}

// size pass of one frame
GOFR_HD PdnRow pdn_decode_size(const PbnDesc& D, const uint8_t* f, uint32_t fn) {
    PdnRow r = {GOFR_GRPC_OK, 0};
    if (fn < 5) { r.status = GOFR_GRPC_BAD_LENGTH; return r; }
    if (f[0] == 1) { r.status = GOFR_GRPC_COMPRESSED; return r; }
    if (f[0] != 0) { r.status = GOFR_GRPC_BAD_LENGTH; return r; }
    const uint32_t L = (uint32_t)f[1] << 24 | (uint32_t)f[2] << 16 | (uint32_t)f[3] << 8 | f[4];
    if (L != fn - 5) { r.status = GOFR_GRPC_BAD_LENGTH; return r; }
    r.status = pdn_validate(D, f, 5, 5 + L);
    if (r.status != GOFR_GRPC_OK) return r;
    r.out_len = (pdn_walk<false>(D, f, 5, 5 + L, nullptr) + 3u) & ~3u;
    return r;
}

// emit pass: the row at dst (4-byte aligned: every row is a multiple of 4 bytes long)
GOFR_HD void pdn_decode_emit(const PbnDesc& D, const uint8_t* f, uint32_t fn, const PdnRow r, uint8_t* dst) {
    if (!r.out_len) return;
    uint32_t end = pdn_walk<true>(D, f, 5, fn, dst);
    for (; end < r.out_len; end++) dst[end] = 0;
}

}  // namespace gofr
