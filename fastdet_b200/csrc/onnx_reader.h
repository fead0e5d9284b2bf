// onnx_reader.h — dependency-free reader for the subset of the ONNX protobuf schema the YOLOv3 graphs
// use.  Replaces the graph-loading half of `ort.InferenceSession(path, providers)` (reference
// server/detector.py:118).  Field numbers are those of the public onnx.proto.
#pragma once
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

namespace fd {

struct OnnxTensor {
    std::string name;
    std::vector<int64_t> dims;
    int dtype = 0;               // 1 f32, 6 i32, 7 i64, 10 f16, 11 f64
    std::vector<float> f;        // floating payload converted to f32
    std::vector<int64_t> i;      // integer payload
    bool is_float() const { return dtype == 1 || dtype == 10 || dtype == 11; }
    size_t numel() const {
        size_t n = 1;
        for (int64_t d : dims) n *= static_cast<size_t>(d);
        return n;
    }
};

struct OnnxAttr {
    int type = 0;  // 1 FLOAT 2 INT 3 STRING 4 TENSOR 6 FLOATS 7 INTS
    float f = 0.f;
    int64_t i = 0;
    std::string s;
    OnnxTensor t;
    std::vector<float> floats;
    std::vector<int64_t> ints;
};

struct OnnxNode {
    std::string op, name;
    std::vector<std::string> inputs, outputs;
    std::map<std::string, OnnxAttr> attrs;
    int64_t attr_i(const char* k, int64_t dflt) const {
        auto it = attrs.find(k);
        return it == attrs.end() ? dflt : it->second.i;
    }
    float attr_f(const char* k, float dflt) const {
        auto it = attrs.find(k);
        return it == attrs.end() ? dflt : it->second.f;
    }
    std::string attr_s(const char* k, const char* dflt) const {
        auto it = attrs.find(k);
        return it == attrs.end() ? std::string(dflt) : it->second.s;
    }
    const std::vector<int64_t>* attr_ints(const char* k) const {
        auto it = attrs.find(k);
        return it == attrs.end() ? nullptr : &it->second.ints;
    }
};

struct OnnxValueInfo {
    std::string name;
    std::vector<int64_t> dims;  // -1 for symbolic
};

struct OnnxGraph {
    std::vector<OnnxNode> nodes;
    std::map<std::string, OnnxTensor> initializers;
    std::vector<OnnxValueInfo> inputs;   // initializers filtered out
    std::vector<OnnxValueInfo> outputs;
    int64_t opset = 0;
    std::string producer;
};

// Returns false and fills `err` on malformed input.
bool onnx_parse(const void* data, size_t len, OnnxGraph* g, std::string* err);

}  // namespace fd
