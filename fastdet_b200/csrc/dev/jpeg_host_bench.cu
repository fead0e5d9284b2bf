// dev tool: single-thread timing of the host half of the JPEG front end (parse + Huffman decode).
//   build/jpeg_host_bench file.jpg [reps]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>

#include "../jpeg.h"

int main(int argc, char** argv) {
    if (argc < 2) return 1;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 1;
    std::vector<uint8_t> d;
    uint8_t buf[65536];
    size_t k;
    while ((k = fread(buf, 1, sizeof(buf), f)) > 0) d.insert(d.end(), buf, buf + k);
    fclose(f);
    const int reps = argc > 2 ? atoi(argv[2]) : 200;
    std::unique_ptr<fd::JpegInfo> J(new fd::JpegInfo());
    char why[160];
    if (fd::jpeg_parse(d.data(), d.size(), J.get(), why, sizeof(why))) { printf("parse: %s\n", why); return 1; }
    std::vector<int16_t> coef(fd::jpeg_coef_count(*J));
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < reps; ++i) fd::jpeg_parse(d.data(), d.size(), J.get(), why, sizeof(why));
    auto t1 = std::chrono::steady_clock::now();
    long long sum = 0;
    for (int i = 0; i < reps; ++i) {
        if (fd::jpeg_decode_coefficients(d.data(), d.size(), *J, coef.data(), why, sizeof(why))) { printf("decode: %s\n", why); return 1; }
        sum += coef[0];
    }
    auto t2 = std::chrono::steady_clock::now();
    size_t nz = 0;
    for (int16_t c : coef) nz += c != 0;
    const double parse_us = std::chrono::duration<double, std::micro>(t1 - t0).count() / reps;
    const double dec_us = std::chrono::duration<double, std::micro>(t2 - t1).count() / reps;
    printf("%zu bytes, %zu nonzero coefficients: parse %.1f us, entropy decode %.1f us (%.1f MB/s, %.1f ns/coefficient) [%lld]\n",
           d.size(), nz, parse_us, dec_us, d.size() / dec_us, dec_us * 1e3 / nz, sum);
    return 0;
}
