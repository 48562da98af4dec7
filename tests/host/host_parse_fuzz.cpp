// Mutation fuzzing of the PLONK / fflonk zkey parsers and of the host-side load logic that consumes their output
// (plonk_parse_zkey, fflonk_parse_zkey, the section reads sb_plonk_load / sb_fflonk_load perform, plonk_addition_levels),
// built with -fsanitize=address,undefined by tests/test_host_templates.py.  Every mutated container must either be
// rejected with a message or be accepted with every section read staying inside the buffer; the sanitizers turn an
// out-of-bounds read or an overflowed size into a crash.  The buffer handed to the parser is an exact-size heap block,
// so reading one byte past the file is detected.   usage: host_parse_fuzz plonk|fflonk <zkey file> <iterations>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iterator>
#include "../../snarkjs_b200/csrc/fflonk_flow.h"
using namespace sb;

static uint64_t st = 0x243F6A8885A308D3ull;
static uint64_t rnd() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return st; }
static volatile uint64_t sink;
static void touch(const uint8_t* p, uint64_t len) { uint64_t s = 0; for (uint64_t i = 0; i < len; i += 97) s += p[i]; if (len) s += p[len - 1]; sink += s; }

// what sb_plonk_load reads after a successful parse (api_plonk.inl plonk_load_impl), with host reads in place of uploads
static void consume_plonk(const PlonkZkey& z) {
    const uint64_t n = z.n, sd = n * 32;
    touch(z.sec[2].p, z.sec[2].len); touch(z.q, z.n8q); touch(z.r, z.n8r); touch(z.k1, 32); touch(z.k2, 32);
    touch(z.hdr_pts, 8 * 2 * (uint64_t)z.n8q); touch(z.X_2, 4 * (uint64_t)z.n8q);
    const uint32_t na = z.nAdditions;
    std::vector<uint32_t> sig(2 * (size_t)na + 2), order, level_end;
    for (uint32_t i = 0; i < na; i++) { memcpy(&sig[2 * (size_t)i], z.sec[3].p + 72 * (size_t)i, 8); touch(z.sec[3].p + 72 * (size_t)i + 8, 64); }
    plonk_addition_levels(sig.data(), na, z.nVars - na, order, level_end);
    uint64_t tot = 0; for (uint32_t e : level_end) tot = e;
    if (na && tot != na) { fprintf(stderr, "addition levels do not cover all additions\n"); abort(); }
    for (uint32_t o : order) if (o >= na) { fprintf(stderr, "addition order out of range\n"); abort(); }
    for (int j = 0; j < 3; j++) touch(z.sec[4 + j].p, (uint64_t)z.nConstraints * 4);
    for (int j = 0; j < 5; j++) touch(z.sec[7 + j].p, 5 * sd);
    touch(z.sec[12].p, 15 * sd);
    const uint32_t nl = z.nPublic > 1 ? z.nPublic : 1, have = (uint32_t)(z.sec[13].len / (5 * sd));
    for (uint32_t j = 0; j < nl && j < have; j++) touch(z.sec[13].p + 5 * sd * j + sd, 4 * sd);
    touch(z.sec[14].p, (n + 6) * 2 * (uint64_t)z.n8q);
}

int main(int argc, char** argv) {
    if (argc < 4) return 2;
    const bool ff = std::string(argv[1]) == "fflonk";
    std::ifstream f(argv[2], std::ios::binary);
    std::vector<uint8_t> good((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    const long iters = atol(argv[3]);
    if (good.size() < 64) { fprintf(stderr, "cannot read %s\n", argv[2]); return 2; }
    long accepted = 0, rejected = 0;
    for (long it = 0; it <= iters; it++) {
        uint64_t len = good.size();
        const int mode = it == 0 ? -1 : (int)(rnd() % 6);
        if (mode == 0) len = rnd() % (good.size() + 1);                       // truncation
        uint8_t* buf = (uint8_t*)malloc(len ? len : 1);                         // exact size: ASan sees any overrun
        memcpy(buf, good.data(), len);
        if (len >= 64) {
            if (mode == 1) for (int k = 0; k < 1 + (int)(rnd() % 4); k++) buf[rnd() % std::min<uint64_t>(len, 2048)] ^= (uint8_t)(1u << (rnd() % 8));   // bit flips in the headers / section table
            if (mode == 2) { uint64_t p = rnd() % std::min<uint64_t>(len - 8, 4096); uint64_t v = rnd() % 3 == 0 ? ~0ull - (rnd() % 64) : (rnd() % 2 ? rnd() : rnd() % (2 * len)); memcpy(buf + p, &v, 8); }   // a wild 64-bit field
            if (mode == 3) { uint64_t p = rnd() % std::min<uint64_t>(len - 4, 4096); uint32_t v = rnd() % 2 ? (uint32_t)rnd() : (uint32_t)(rnd() % 64); memcpy(buf + p, &v, 4); }       // a wild 32-bit field (counts, sizes)
            if (mode == 4) for (int k = 0; k < 16; k++) { uint64_t p = rnd() % (len - 4); uint32_t v = (uint32_t)rnd(); memcpy(buf + p, &v, 4); }                                      // garbage anywhere (signal ids, maps)
            if (mode == 5) { uint64_t p = 12 + rnd() % std::min<uint64_t>(len - 12, 600); memset(buf + p, rnd() % 2 ? 0xff : 0, std::min<uint64_t>(12, len - p)); }                   // a smashed section header
        }
        std::string err; int rc;
        if (ff) { FflonkZkey z; rc = fflonk_parse_zkey(buf, len, z, err); if (!rc) { for (int id = 1; id < 20; id++) if (z.sec[id].p) touch(z.sec[id].p, z.sec[id].len); touch(z.q, z.n8q); touch(z.r, z.n8r); if (z.C0) touch(z.C0, 2 * (uint64_t)z.n8q); } }
        else { PlonkZkey z; rc = plonk_parse_zkey(buf, len, z, err); if (!rc) consume_plonk(z); }
        if (rc) { if (err.empty()) { fprintf(stderr, "rejected without a message (iteration %ld)\n", it); return 1; } rejected++; }
        else accepted++;
        if (it == 0 && rc) { fprintf(stderr, "the unmodified key was rejected: %s\n", err.c_str()); return 1; }
        free(buf);
    }
    printf("FUZZ OK %s: %ld accepted, %ld rejected\n", argv[1], accepted, rejected);
    return 0;
}
