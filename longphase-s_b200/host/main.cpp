// main.cpp — `longphase-s-b200 <command> [options]`: the reference's sub-command dispatch (src/main.cpp:29-64) for the
// sub-commands whose hot path runs through liblps_b200.so.
#include <iostream>
#include <string>

#include "lps_host.h"

static const char *USAGE =
    "Usage: longphase-s-b200 <command> [options]\n"
    "    phase                  run phasing algorithm (GPU hot path).\n"
    "    haplotag               tag reads by haplotype (GPU hot path).\n"
    "    somatic_haplotag       tag tumor reads by somatic haplotype: extract passes, purity, calling, tagging (GPU hot path).\n"
    "    estimate_purity        tumor purity from a tumor / normal pair: the two extract passes + the estimate (GPU hot path).\n\n";

int main(int argc, char **argv) {
    if (argc <= 1) { std::cout << USAGE; return 0; }
    const std::string command(argv[1]);
    if (command == "phase") return lpsh_phase_main(argc - 1, argv + 1);
    if (command == "haplotag") return lpsh_tag_main(argc - 1, argv + 1);
    if (command == "somatic_haplotag" || command == "estimate_purity") return lpsh_som_main(argc - 1, argv + 1);
    std::cout << USAGE;
    return 0;
}
