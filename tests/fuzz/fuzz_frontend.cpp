// Test infrastructure: mutation fuzzer for the C++ rule-language front end (YAML reader, parser, GLSL emitters,
// typed expression compiler / CUDA emitter).  Built with ASan + UBSan by tests/test_frontend_fuzz.py.  Every input must end
// in success or in a se::ParseError -- never in a crash, a sanitizer report or a foreign exception.
// usage: fuzz_frontend <iterations> <seed yaml>...
#include "lang/lang.h"
#include "lang/codegen.h"
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <random>
#include <sstream>
#include <string>
int main(int argc, char** argv) {
    std::vector<std::string> seeds;
    for (int i = 2; i < argc; ++i) { std::ifstream f(argv[i]); std::stringstream ss; ss << f.rdbuf(); seeds.push_back(ss.str()); }
    int n_iter = std::atoi(argv[1]);
    const char* seed_env = std::getenv("FUZZ_SEED");          // default: the fixed seed the test suite uses
    std::mt19937 rng(seed_env ? (unsigned)std::strtoul(seed_env, nullptr, 10) : 12345u);
    const char alphabet[] = " \n:-[]{},#'\"!|&<>=.()_abcdefSELFDOWNRIGHTLEFTmat0123456789\t%*+/";
    long ok = 0, perr = 0, other = 0;
    for (int it = 0; it < n_iter; ++it) {
        std::string s = seeds[rng() % seeds.size()];
        int nm = 1 + rng() % 4;
        for (int m = 0; m < nm && !s.empty(); ++m) {
            size_t p = rng() % s.size();
            switch (rng() % 5) {
                case 0: s.erase(p, 1 + rng() % 6); break;
                case 1: s.insert(p, 1, alphabet[rng() % (sizeof alphabet - 1)]); break;
                case 2: s[p] = alphabet[rng() % (sizeof alphabet - 1)]; break;
                case 3: { size_t q = rng() % s.size(); size_t n = 1 + rng() % 20; s.insert(p, s.substr(q, n)); break; }
                case 4: { size_t e = s.find('\n', p); if (e != std::string::npos) s.erase(p, e - p); break; }
            }
        }
        try {
            se::ParsingResult r = se::parse_string(s);
            se::emit_glsl_materials(r); se::emit_glsl_rules(r);
            se::CompiledRules c = se::compile_rules(r);
            ++ok;
            if (const char* dump = std::getenv("FUZZ_DUMP_DIR")) {     // keep accepted mutants (e.g. to push them through NVRTC)
                std::ofstream(std::string(dump) + "/ok_" + std::to_string(it) + ".yaml") << s;
            }
        } catch (const se::ParseError&) { ++perr; }
        catch (const std::exception& e) { ++other; if (other < 10) std::printf("other exception: %s\n", e.what()); }
    }
    std::printf("ok=%ld parse_errors=%ld other=%ld\n", ok, perr, other);
    return 0;
}
