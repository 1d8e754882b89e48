// Minimal stand-alone runner of the hot path with the reference CLI's shape:
//   b200_pipeline <pipeline> baseband <input> <output_hint> --samplerate 6e6 --baseband_format cs16 [--fused] [--key value ...]
// (cf. `satdump pipeline metop_ahrpt baseband in.cs16 out --samplerate 6e6 --baseband_format cs16`, README.md:52).
// Pipelines known here carry the module parameters of resources/pipelines/{MetOp,JPSS}.json.
#include "stream_modules.hpp"
#include <cstdio>
#include <cstring>
using namespace b200host;

int main(int argc, char **argv)
{
    if (argc < 5) {
        fprintf(stderr, "usage: %s <metop_ahrpt|jpss_hrd|npp_hrd|simple_bpsk|simple_qpsk|pm_bpsk> baseband <input> <output_hint> [--fused] [--key value ...]\n", argv[0]);
        return 2;
    }
    const std::string pipe = argv[1], level = argv[2], in = argv[3], out = argv[4];
    Params dp, fp;
    std::string dec;
    if (pipe == "metop_ahrpt") { // resources/pipelines/MetOp.json:30-47
        dp = Params{{"constellation", "qpsk"}, {"symbolrate", "2333333"}, {"rrc_alpha", "0.5"}, {"pll_bw", "0.003"}};
        fp = Params{{"viterbi_outsync_after", "10"}, {"viterbi_ber_thresold", "0.28"}};
        dec = "metop_ahrpt_decoder";
    } else if (pipe == "jpss_hrd" || pipe == "npp_hrd") { // resources/pipelines/JPSS.json
        dp = Params{{"constellation", "oqpsk"}, {"symbolrate", pipe == "jpss_hrd" ? "25000000" : "15000000"}, {"rrc_alpha", "0.5"}, {"pll_bw", "0.002"}};
        fp = Params{{"constellation", "oqpsk"}, {"cadu_size", "10232"}, {"viterbi_ber_thresold", "0.3"}, {"viterbi_outsync_after", "20"},
                    {"derandomize", "true"}, {"nrzm", "true"}, {"rs_i", "5"}, {"rs_type", "rs223"}, {"rs_usecheck", "true"}};
        dec = "ccsds_conv_concat_decoder";
    } else if (pipe == "simple_bpsk" || pipe == "simple_qpsk") { // psk_demod -> ccsds_simple_psk_decoder (no convolutional code), the shape of
        // 50 shipped pipelines (e.g. resources/pipelines/FengYun-3.json:630-639); symbolrate etc. come from the command line
        const char *con = pipe == "simple_bpsk" ? "bpsk" : "qpsk";
        dp = Params{{"constellation", con}, {"symbolrate", "1200000"}, {"rrc_alpha", "0.5"}, {"pll_bw", "0.003"}};
        fp = Params{{"constellation", con}, {"cadu_size", "8192"}, {"nrzm", pipe == "simple_bpsk" ? "true" : "false"}, {"derandomize", "true"}, {"rs_i", "4"},
                    {"rs_type", "rs223"}};
        dec = "ccsds_simple_psk_decoder";
    } else if (pipe == "pm_bpsk") { // pm_demod -> ccsds_conv_concat_decoder, the shape of 14 shipped pipelines (residual-carrier PM with a BPSK
        // subcarrier; symbolrate etc. come from the command line)
        dp = Params{{"b200_demod_module", "pm_demod"}, {"symbolrate", "500000"}, {"rrc_alpha", "0.5"}, {"pll_bw", "0.01"}, {"pll_max_offset", "3.14"},
                    {"costas_bw", "0.005"}};
        fp = Params{{"constellation", "bpsk"}, {"cadu_size", "8192"}, {"viterbi_ber_thresold", "0.3"}, {"viterbi_outsync_after", "20"}, {"derandomize", "true"},
                    {"rs_i", "4"}, {"rs_type", "rs223"}};
        dec = "ccsds_conv_concat_decoder";
    } else {
        fprintf(stderr, "unknown pipeline %s\n", pipe.c_str());
        return 2;
    }
    bool fused = false;
    for (int i = 5; i < argc; i++) {
        if (!strcmp(argv[i], "--fused")) { fused = true; continue; }
        if (!strncmp(argv[i], "--", 2) && i + 1 < argc) { dp.set(argv[i] + 2, argv[i + 1]); fp.set(argv[i] + 2, argv[i + 1]); i++; }
    }
    if (level != "baseband") {
        fprintf(stderr, "only the 'baseband' input level is part of the hot path\n");
        return 2;
    }
    try {
        if (fused) {
            FusedStage m(dec, in, out, dp, fp);
            m.process();
            printf("wrote %s (%ld frames)\n", m.getOutput().c_str(), m.frames_written.load());
        } else {
            PskDemodStage m1(in, out, dp, dp.str("b200_demod_module", "psk_demod"));
            FecStage m2(dec, "", out, fp);
            run_two_stage(m1, m2);
            printf("wrote %s (%ld frames)\n", m2.getOutput().c_str(), m2.frames_written.load());
        }
    } catch (const std::exception &e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
