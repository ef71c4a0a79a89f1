// proto_nested_device.cuh — proto3 encoder for message types with nested and repeated fields
// (gofr_proto_encode_nested_device): proto.Marshal of the message a unary handler returns + grpc-go's 5-byte
// length-prefixed-message header (examples/grpc-server/grpc/hello_grpc.pb.go:73-89 hands the message to grpc-go's proto
// codec; protobuf-go v1.32.0, grpc-go v1.60.1) — the widening of grpc_device.cuh's flat encoder that SURVEY.md §8f rank 4
// leaves open.  Wire rules (protobuf-go impl/encode.go, codec_field.go, codec_gen.go):
//   * fields in ascending field-number order; a singular scalar holding its zero value is not written (proto3 implicit
//     presence); strings must be valid UTF-8;
//   * a singular message field is written when it is set, even when the message is empty (tag, length 0);
//   * repeated scalars of numeric type are PACKED: one tag (wire type 2), the payload length, the elements back to back —
//     nothing at all for zero elements; repeated strings / bytes / messages: one tag + length + payload PER ELEMENT,
//     empty elements included;
//   * map fields are not modelled: protobuf-go writes them in Go's random map order unless Deterministic is set, so there
//     is no byte sequence to be identical to.
// Rows use the layout of the wider data model (include/gofr_b200.h "Row format"): fixed words of a message in field order
// (64-bit kinds two words, string / bytes their length, a singular message a presence word + its own fixed part inline, a
// repeated field its element count), then the variable part in field order (string bytes; a set message's variable part;
// repeated elements: scalars as their words, strings as u32 length + bytes, messages as fixed part + variable part).
// Device code does not recurse: the walk keeps an explicit stack of at most kPbnMaxDepth message levels (checked on the
// host when the call is made); the length prefix of a nested message comes from a sizing walk of that message.
// __host__ __device__ like the rest: tests/emu runs it on the CPU against the oracle and python google.protobuf.
#pragma once
#include "grpc_device.cuh"

namespace gofr {

struct PbnFrame {
    const uint8_t* fixed;  // fixed part of the message this frame walks
    uint32_t size;         // SIZE walk: bytes of the message's content so far
    uint32_t rep_left;     // elements still to come of the repeated message field the frame is in the middle of
    uint8_t mt, fi;        // message type, next field (index inside the type)
    uint8_t pend_tl;       // SIZE walk: tag length of the message field whose content the child frame is sizing
};

GOFR_HD uint32_t pbn_ld32(const uint8_t* p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }

// One walk over message type `mt` whose fixed part is at `fixed` and whose variable part starts at *var (end = end of the
// row).  EMIT == false: returns the size of the message's content, validates the row (*status: GOFR_GRPC_BAD_ROW /
// GOFR_GRPC_BAD_UTF8) and advances *var past the message's variable part.  EMIT == true: writes the content through w
// (the row has been validated by the sizing walk of the whole row).
template <bool EMIT>
GOFR_HD_NOINLINE uint32_t pbn_walk(const PbnDesc& D, uint32_t mt, const uint8_t* fixed, const uint8_t** var_io, const uint8_t* end,
                                   Writer* w, uint32_t* status) {
    PbnFrame st[kPbnMaxDepth + 1];
    int depth = 0;
    const uint8_t* var = *var_io;
    uint32_t err = GOFR_GRPC_OK, result = 0;
    auto take = [&](uint32_t n) -> const uint8_t* {
        if ((uint32_t)(end - var) < n) { err = GOFR_GRPC_BAD_ROW; return nullptr; }
        const uint8_t* p = var;
        var += n;
        return p;
    };
    auto push = [&](uint32_t m, const uint8_t* fx) {
        if (depth > kPbnMaxDepth) { err = GOFR_GRPC_BAD_ROW; return; }
        PbnFrame& f = st[depth++];
        f.fixed = fx; f.size = 0; f.rep_left = 0; f.mt = (uint8_t)m; f.fi = 0; f.pend_tl = 0;
    };
    auto put_varint = [&](uint64_t v) { w->reserve(3); proto_put_varint(*w, v); };
    // a scalar payload (what follows the tag) of class cls whose words are at p: its encoded length, written when EMIT
    auto scalar = [&](uint32_t cls, const uint8_t* p) -> uint32_t {
        const uint32_t w0 = pbn_ld32(p), w1 = (cls & PC_64) ? pbn_ld32(p + 4) : 0u;
        const uint32_t wire = cls & PC_WIRE;
        if (wire == 1) { if (EMIT) { w->reserve(3); w->put4(w0); w->put4(w1); } return 8; }
        if (wire == 5) { if (EMIT) { w->reserve(2); w->put4(w0); } return 4; }
        const uint64_t v = proto_varint_value(cls, w0, w1);
        if (EMIT) put_varint(v);
        return varint_len64(v);
    };
    // enters the message of field F (type F.msg) whose fixed part is at fx: the tag and length prefix now (EMIT) or when
    // the child frame is popped (SIZE)
    auto enter_message = [&](const PbnField& F, const uint8_t* fx, int parent) {
        if constexpr (EMIT) {  // (constexpr: the sizing instance must not contain a call to itself)
            const uint8_t* v2 = var;
            uint32_t st2 = GOFR_GRPC_OK;
            const uint32_t sub = pbn_walk<false>(D, F.msg, fx, &v2, end, nullptr, &st2);
            put_varint(F.tag);
            put_varint(sub);
        } else st[parent].pend_tl = (uint8_t)varint_len(F.tag);
        push(F.msg, fx);
    };

    push(mt, fixed);
    while (depth > 0 && err == GOFR_GRPC_OK) {
        const int fi = depth - 1;
        const PbnField* fields = D.f + D.first[st[fi].mt];
        if (st[fi].rep_left) {  // next element of a repeated message field (the field before st[fi].fi)
            const PbnField& F = fields[st[fi].fi - 1];
            st[fi].rep_left--;
            const uint8_t* fx = take((uint32_t)D.fixed_words[F.msg] * 4u);
            if (fx) enter_message(F, fx, fi);
            continue;
        }
        if (st[fi].fi >= D.count[st[fi].mt]) {  // message done
            const uint32_t sub = st[fi].size;
            depth--;
            if (depth > 0) { if (!EMIT) st[depth - 1].size += st[depth - 1].pend_tl + varint_len(sub) + sub; }
            else result = sub;
            continue;
        }
        const PbnField& F = fields[st[fi].fi];
        // the field's fixed words: fields before it in the same message
        uint32_t wo = 0;
        for (uint32_t k = 0; k < st[fi].fi; k++) wo += fields[k].fixed_words;
        const uint8_t* p = st[fi].fixed + (size_t)wo * 4;
        st[fi].fi++;
        const uint32_t tl = varint_len(F.tag), cls = F.cls, wire = cls & PC_WIRE;
        if (!F.repeated) {
            if (F.msg != 0xFF) {  // singular message: presence word, then the fixed part inline
                if (pbn_ld32(p)) enter_message(F, p + 4, fi);
            } else if (wire == 2) {  // string / bytes
                const uint32_t len = pbn_ld32(p);
                const uint8_t* s = take(len);
                if (!s) break;
                if (!EMIT && (cls & PC_UTF8) && !proto_utf8_ok(s, len)) { err = GOFR_GRPC_BAD_UTF8; break; }
                if (len) {
                    if (EMIT) { put_varint(F.tag); put_varint(len); w->copy<false>(s, len); }
                    st[fi].size += tl + varint_len(len) + len;
                }
            } else {
                const uint32_t w0 = pbn_ld32(p), w1 = (cls & PC_64) ? pbn_ld32(p + 4) : 0u;
                if (w0 | w1) {
                    if (EMIT) put_varint(F.tag);
                    st[fi].size += tl + scalar(cls, p);
                }
            }
            continue;
        }
        const uint32_t n = pbn_ld32(p);
        if (F.msg != 0xFF) {  // repeated messages: one after the other through the frame's rep_left
            if ((uint32_t)(end - var) / 4u < n && D.fixed_words[F.msg]) { err = GOFR_GRPC_BAD_ROW; break; }
            if (n > 0x00FFFFFFu) { err = GOFR_GRPC_BAD_ROW; break; }
            st[fi].rep_left = n;
        } else if (wire == 2) {  // repeated string / bytes: tag + length + payload per element
            if ((uint32_t)(end - var) / 4u < n) { err = GOFR_GRPC_BAD_ROW; break; }
            for (uint32_t i = 0; i < n; i++) {
                const uint8_t* lp = take(4);
                if (!lp) break;
                const uint32_t len = pbn_ld32(lp);
                const uint8_t* s = take(len);
                if (!s) break;
                if (!EMIT && (cls & PC_UTF8) && !proto_utf8_ok(s, len)) { err = GOFR_GRPC_BAD_UTF8; break; }
                if (EMIT) { put_varint(F.tag); put_varint(len); if (len) w->copy<false>(s, len); }
                st[fi].size += tl + varint_len(len) + len;
            }
        } else if (n) {  // packed repeated scalars: tag, payload length, the elements
            const uint32_t eb = (cls & PC_64) ? 8u : 4u;
            if ((uint32_t)(end - var) / eb < n) { err = GOFR_GRPC_BAD_ROW; break; }
            const uint8_t* e0 = take(n * eb);
            uint32_t payload = 0;
            if (wire == 1) payload = 8u * n;
            else if (wire == 5) payload = 4u * n;
            else for (uint32_t i = 0; i < n; i++) {
                const uint8_t* q = e0 + (size_t)i * eb;
                payload += varint_len64(proto_varint_value(cls, pbn_ld32(q), (cls & PC_64) ? pbn_ld32(q + 4) : 0u));
            }
            if (EMIT) {
                put_varint(F.tag);
                put_varint(payload);
                for (uint32_t i = 0; i < n; i++) scalar(cls, e0 + (size_t)i * eb);
            }
            st[fi].size += tl + varint_len(payload) + payload;
        }
    }
    *var_io = var;
    *status = err;
    return result;
}

struct PbnMsg {
    uint32_t status;   // GOFR_GRPC_OK / GOFR_GRPC_BAD_UTF8 / GOFR_GRPC_BAD_ROW
    uint32_t out_len;  // 5 + message bytes (0 on error)
};

// size pass of one row: validates it and returns the exact frame length
GOFR_HD PbnMsg pbn_size(const PbnDesc& D, const uint8_t* row, uint32_t rn) {
    PbnMsg m = {GOFR_GRPC_OK, 0};
    const uint32_t fb = (uint32_t)D.fixed_words[D.root] * 4u;
    if (rn < fb) { m.status = GOFR_GRPC_BAD_ROW; return m; }
    const uint8_t* var = row + fb;
    uint32_t st = GOFR_GRPC_OK;
    const uint32_t len = pbn_walk<false>(D, D.root, row, &var, row + rn, nullptr, &st);
    m.status = st;
    m.out_len = st == GOFR_GRPC_OK ? 5u + len : 0u;
    return m;
}

// emit pass: the frame at dst (arbitrary alignment inside the packed output)
GOFR_HD void pbn_emit(const PbnDesc& D, const uint8_t* row, uint32_t rn, const PbnMsg m, uint8_t* dst, uint32_t* stage_col) {
    if (!m.out_len) return;
    Writer w;
    w.init(dst, stage_col);
    const uint32_t plen = m.out_len - 5;
    w.put4(0u | (plen >> 24) << 8 | ((plen >> 16) & 0xFF) << 16 | ((plen >> 8) & 0xFF) << 24);  // 00, be32[0..2]
    w.putc(plen & 0xFF);
    const uint8_t* var = row + (size_t)D.fixed_words[D.root] * 4;
    uint32_t st = GOFR_GRPC_OK;
    pbn_walk<true>(D, D.root, row, &var, row + rn, &w, &st);
    w.finish();
}

}  // namespace gofr
