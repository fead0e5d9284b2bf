// onnx_reader.cc — see onnx_reader.h.  A bounds-checked protobuf wire-format walker; repeated scalar
// fields are accepted both packed and unpacked (torch's serializer writes unpacked ints, the onnx
// package writes packed ones).
#include "onnx_reader.h"

#include <math.h>
#include <string.h>

namespace fd {
namespace {

struct Pb {
    const uint8_t* p;
    const uint8_t* end;
    bool ok = true;

    Pb(const void* d, size_t n) : p(static_cast<const uint8_t*>(d)), end(static_cast<const uint8_t*>(d) + n) {}
    bool done() const { return p >= end || !ok; }
    uint64_t varint() {
        uint64_t v = 0;
        int shift = 0;
        while (p < end && shift < 64) {
            const uint8_t c = *p++;
            v |= static_cast<uint64_t>(c & 0x7F) << shift;
            if (!(c & 0x80)) return v;
            shift += 7;
        }
        ok = false;
        return 0;
    }
    // next field header; for wire type 2 `sub` receives the payload span
    bool next(uint32_t* field, uint32_t* wt, uint64_t* scalar, Pb* sub) {
        if (done()) return false;
        const uint64_t key = varint();
        if (!ok) return false;
        *field = static_cast<uint32_t>(key >> 3);
        *wt = static_cast<uint32_t>(key & 7);
        switch (*wt) {
            case 0: *scalar = varint(); return ok;
            case 1:
                if (end - p < 8) { ok = false; return false; }
                memcpy(scalar, p, 8); p += 8; return true;
            case 5: {
                if (end - p < 4) { ok = false; return false; }
                uint32_t v; memcpy(&v, p, 4); p += 4; *scalar = v; return true;
            }
            case 2: {
                const uint64_t n = varint();
                if (!ok || n > static_cast<uint64_t>(end - p)) { ok = false; return false; }
                *sub = Pb(p, static_cast<size_t>(n));
                p += n;
                return true;
            }
            default: ok = false; return false;
        }
    }
    std::string str() const { return std::string(reinterpret_cast<const char*>(p), static_cast<size_t>(end - p)); }
    size_t size() const { return static_cast<size_t>(end - p); }
};

float half_to_float(uint16_t h) {
    const uint32_t sign = (h >> 15) & 1, exp = (h >> 10) & 0x1F, man = h & 0x3FF;
    float v;
    if (exp == 0) v = ldexpf(static_cast<float>(man), -24);
    else if (exp == 31) v = man ? NAN : INFINITY;
    else v = ldexpf(static_cast<float>(man | 0x400), static_cast<int>(exp) - 25);
    return sign ? -v : v;
}

void read_packed_or_single_i64(uint32_t wt, uint64_t scalar, Pb sub, std::vector<int64_t>* out) {
    if (wt == 2) {
        while (!sub.done()) {
            const uint64_t v = sub.varint();
            if (sub.ok) out->push_back(static_cast<int64_t>(v));
        }
    } else {
        out->push_back(static_cast<int64_t>(scalar));
    }
}

void read_packed_or_single_f32(uint32_t wt, uint64_t scalar, Pb sub, std::vector<float>* out) {
    if (wt == 2) {
        const size_t n = sub.size() / 4;
        const size_t base = out->size();
        out->resize(base + n);
        memcpy(out->data() + base, sub.p, n * 4);
    } else {
        const uint32_t u = static_cast<uint32_t>(scalar);
        float f;
        memcpy(&f, &u, 4);
        out->push_back(f);
    }
}

bool parse_tensor(Pb pb, OnnxTensor* t, std::string* err) {
    uint32_t f, wt;
    uint64_t sc = 0;
    Pb sub(nullptr, 0);
    Pb raw(nullptr, 0);
    bool has_raw = false;
    std::vector<float> floats;
    std::vector<int64_t> ints;
    std::vector<double> doubles;
    while (pb.next(&f, &wt, &sc, &sub)) {
        switch (f) {
            case 1: read_packed_or_single_i64(wt, sc, sub, &t->dims); break;
            case 2: t->dtype = static_cast<int>(sc); break;
            case 4: read_packed_or_single_f32(wt, sc, sub, &floats); break;
            case 5:
            case 7: read_packed_or_single_i64(wt, sc, sub, &ints); break;
            case 8: if (wt == 2) t->name = sub.str(); break;
            case 9: if (wt == 2) { raw = sub; has_raw = true; } break;
            case 10:
                if (wt == 2) {
                    const size_t n = sub.size() / 8;
                    const size_t base = doubles.size();
                    doubles.resize(base + n);
                    memcpy(doubles.data() + base, sub.p, n * 8);
                } else if (wt == 1) {
                    double d;
                    memcpy(&d, &sc, 8);
                    doubles.push_back(d);
                }
                break;
            default: break;
        }
    }
    if (!pb.ok) { *err = "malformed TensorProto"; return false; }
    const size_t n = t->numel();
    switch (t->dtype) {
        case 1:
            if (has_raw) {
                if (raw.size() != n * 4) { *err = "tensor '" + t->name + "': raw_data size mismatch"; return false; }
                t->f.resize(n);
                memcpy(t->f.data(), raw.p, n * 4);
            } else {
                t->f.swap(floats);
            }
            break;
        case 11:
            if (has_raw) {
                if (raw.size() != n * 8) { *err = "tensor '" + t->name + "': raw_data size mismatch"; return false; }
                t->f.resize(n);
                for (size_t k = 0; k < n; ++k) { double d; memcpy(&d, raw.p + 8 * k, 8); t->f[k] = static_cast<float>(d); }
            } else {
                t->f.assign(doubles.begin(), doubles.end());
            }
            break;
        case 10:
            t->f.resize(n);
            if (has_raw) {
                if (raw.size() != n * 2) { *err = "tensor '" + t->name + "': raw_data size mismatch"; return false; }
                for (size_t k = 0; k < n; ++k) { uint16_t h; memcpy(&h, raw.p + 2 * k, 2); t->f[k] = half_to_float(h); }
            } else {
                if (ints.size() != n) { *err = "tensor '" + t->name + "': f16 payload size mismatch"; return false; }
                for (size_t k = 0; k < n; ++k) t->f[k] = half_to_float(static_cast<uint16_t>(ints[k]));
            }
            break;
        case 6:
            if (has_raw) {
                if (raw.size() != n * 4) { *err = "tensor '" + t->name + "': raw_data size mismatch"; return false; }
                t->i.resize(n);
                for (size_t k = 0; k < n; ++k) { int32_t v; memcpy(&v, raw.p + 4 * k, 4); t->i[k] = v; }
            } else {
                t->i.resize(ints.size());
                for (size_t k = 0; k < ints.size(); ++k) t->i[k] = static_cast<int32_t>(ints[k]);
            }
            break;
        case 7:
            if (has_raw) {
                if (raw.size() != n * 8) { *err = "tensor '" + t->name + "': raw_data size mismatch"; return false; }
                t->i.resize(n);
                memcpy(t->i.data(), raw.p, n * 8);
            } else {
                t->i.swap(ints);
            }
            break;
        default:
            *err = "tensor '" + t->name + "': unsupported data_type " + std::to_string(t->dtype);
            return false;
    }
    const size_t got = t->is_float() ? t->f.size() : t->i.size();
    if (got != n) { *err = "tensor '" + t->name + "': element count mismatch"; return false; }
    return true;
}

bool parse_attr(Pb pb, std::string* name, OnnxAttr* a, std::string* err) {
    uint32_t f, wt;
    uint64_t sc = 0;
    Pb sub(nullptr, 0);
    bool has_f = false, has_i = false, has_s = false, has_t = false;
    while (pb.next(&f, &wt, &sc, &sub)) {
        switch (f) {
            case 1: if (wt == 2) *name = sub.str(); break;
            case 2: { uint32_t u = static_cast<uint32_t>(sc); memcpy(&a->f, &u, 4); has_f = true; break; }
            case 3: a->i = static_cast<int64_t>(sc); has_i = true; break;
            case 4: if (wt == 2) { a->s = sub.str(); has_s = true; } break;
            case 5: if (wt == 2) { if (!parse_tensor(sub, &a->t, err)) return false; has_t = true; } break;
            case 7: read_packed_or_single_f32(wt, sc, sub, &a->floats); break;
            case 8: read_packed_or_single_i64(wt, sc, sub, &a->ints); break;
            case 20: a->type = static_cast<int>(sc); break;
            default: break;
        }
    }
    if (!pb.ok) { *err = "malformed AttributeProto"; return false; }
    if (a->type == 0) {  // IR < 3 writers omit `type`
        if (has_t) a->type = 4; else if (!a->ints.empty()) a->type = 7; else if (!a->floats.empty()) a->type = 6;
        else if (has_s) a->type = 3; else if (has_i) a->type = 2; else if (has_f) a->type = 1;
    }
    return true;
}

bool parse_node(Pb pb, OnnxNode* n, std::string* err) {
    uint32_t f, wt;
    uint64_t sc = 0;
    Pb sub(nullptr, 0);
    while (pb.next(&f, &wt, &sc, &sub)) {
        if (wt != 2) continue;
        switch (f) {
            case 1: n->inputs.push_back(sub.str()); break;
            case 2: n->outputs.push_back(sub.str()); break;
            case 3: n->name = sub.str(); break;
            case 4: n->op = sub.str(); break;
            case 5: {
                std::string an;
                OnnxAttr a;
                if (!parse_attr(sub, &an, &a, err)) return false;
                n->attrs[an] = std::move(a);
                break;
            }
            default: break;
        }
    }
    if (!pb.ok) { *err = "malformed NodeProto"; return false; }
    return true;
}

void parse_value_info(Pb pb, OnnxValueInfo* vi) {
    uint32_t f, wt;
    uint64_t sc = 0;
    Pb sub(nullptr, 0);
    while (pb.next(&f, &wt, &sc, &sub)) {
        if (f == 1 && wt == 2) vi->name = sub.str();
        if (f == 2 && wt == 2) {  // TypeProto
            Pb tp = sub, s2(nullptr, 0);
            uint32_t f2, w2; uint64_t sc2 = 0;
            while (tp.next(&f2, &w2, &sc2, &s2)) {
                if (f2 != 1 || w2 != 2) continue;  // tensor_type
                Pb tt = s2, s3(nullptr, 0);
                uint32_t f3, w3; uint64_t sc3 = 0;
                while (tt.next(&f3, &w3, &sc3, &s3)) {
                    if (f3 != 2 || w3 != 2) continue;  // shape
                    Pb sh = s3, s4(nullptr, 0);
                    uint32_t f4, w4; uint64_t sc4 = 0;
                    while (sh.next(&f4, &w4, &sc4, &s4)) {
                        if (f4 != 1 || w4 != 2) continue;  // dim
                        Pb dm = s4, s5(nullptr, 0);
                        uint32_t f5, w5; uint64_t sc5 = 0;
                        int64_t v = -1;
                        while (dm.next(&f5, &w5, &sc5, &s5))
                            if (f5 == 1 && w5 == 0) v = static_cast<int64_t>(sc5);
                        vi->dims.push_back(v);
                    }
                }
            }
        }
    }
}

}  // namespace

bool onnx_parse(const void* data, size_t len, OnnxGraph* g, std::string* err) {
    Pb pb(data, len);
    uint32_t f, wt;
    uint64_t sc = 0;
    Pb sub(nullptr, 0), graph(nullptr, 0);
    bool has_graph = false;
    while (pb.next(&f, &wt, &sc, &sub)) {
        if (f == 7 && wt == 2) { graph = sub; has_graph = true; }
        else if (f == 2 && wt == 2) g->producer = sub.str();
        else if (f == 8 && wt == 2) {
            Pb os = sub, s2(nullptr, 0);
            uint32_t f2, w2; uint64_t sc2 = 0;
            std::string domain; int64_t ver = 0;
            while (os.next(&f2, &w2, &sc2, &s2)) {
                if (f2 == 1 && w2 == 2) domain = s2.str();
                if (f2 == 2 && w2 == 0) ver = static_cast<int64_t>(sc2);
            }
            if (domain.empty() || domain == "ai.onnx") g->opset = ver;
        }
    }
    if (!pb.ok || !has_graph) { *err = "not an ONNX ModelProto (no graph)"; return false; }
    std::vector<OnnxValueInfo> inputs;
    while (graph.next(&f, &wt, &sc, &sub)) {
        if (wt != 2) continue;
        switch (f) {
            case 1: {
                OnnxNode n;
                if (!parse_node(sub, &n, err)) return false;
                g->nodes.push_back(std::move(n));
                break;
            }
            case 5: {
                OnnxTensor t;
                if (!parse_tensor(sub, &t, err)) return false;
                std::string nm = t.name;
                g->initializers[nm] = std::move(t);
                break;
            }
            case 11: { OnnxValueInfo vi; parse_value_info(sub, &vi); inputs.push_back(std::move(vi)); break; }
            case 12: { OnnxValueInfo vi; parse_value_info(sub, &vi); g->outputs.push_back(std::move(vi)); break; }
            default: break;
        }
    }
    if (!graph.ok) { *err = "malformed GraphProto"; return false; }
    for (auto& vi : inputs)
        if (!g->initializers.count(vi.name)) g->inputs.push_back(vi);
    return true;
}

}  // namespace fd
