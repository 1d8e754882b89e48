// Host side of the decoder: owns the soft FIFO, the lock machine and the deframer/RS bookkeeping of ONE stream and
// launches the kernels of fec.cuh. Mirrors MetOpAHRPTDecoderModule::process
// (plugins/noaa_metop_support/metop/module_metop_ahrpt_decoder.cpp:34-90) and CCSDSConvConcatDecoderModule::process
// (src-core/pipeline/modules/ccsds/module_ccsds_conv_concat_decoder.cpp:140-200).
#pragma once
#include "fec.cuh"
#include "host_common.h"
#include <vector>

namespace b200
{

struct DefrEvent { long pos; int state; int pad; };

class Fec
{
  public:
    explicit Fec(const b200_fec_cfg &cfg);
    ~Fec();
    // soft FIFO: the producer writes `n` bytes at append_ptr() then calls commit(n)
    int8_t *append_ptr() { return softbuf.p + soft_have; }
    long append_room() const { return (long)softbuf.n - soft_have; }
    void commit(long n) { soft_have += n; total_soft += n; }
    void push_host(const int8_t *h, long n);
    void push_device(const int8_t *d, long n);
    void process(); // decode every complete chunk in the FIFO
    void process_simple(); // kind B200_FEC_SIMPLE: hard decisions -> deframer(s) -> RS
    long pull(uint8_t *host_out, long cap);
    void stats(b200_fec_stats *o);

    b200_fec_cfg cfg;
    VitGeom geom;
    int cadu_bytes, nphases, ph0, ph1, nswap, st_synced, spec_steps;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[3];
    float t_vit = 0, t_frames = 0, t_vit_main = 0; // ms of the last process(): whole Viterbi stage, deframe+RS, k_vit_main alone
    cudaEvent_t evm[2], evf[2];
    long last_main_chunks = 0;
    void reset();

    // device storage
    DevBuf<int8_t> softbuf;
    DevBuf<uint2> dec;
    DevBuf<uint2> idle_dec;
    DevBuf<uint32_t> chunk_bits, fifo;
    DevBuf<int> start_state, rs_err, counters, frame_dst; // frame_dst: destination index of every frame under rs_usecheck
    DevBuf<VitRec> rec;
    DevBuf<TbEdge> tb_edges;
    DevBuf<int> tb_list; // [0] count of chunks redone serially in the last launch, [1..] their indices
    DevBuf<int> redo_list; // chunks decoded again because their speculated start state was wrong (+ the chunks whose BER check follows)
    std::vector<int> h_redo, h_redo_ber, h_start;
    // Viterbi_Depunc (conv_rate 2/3 ... 7/8): the depunctured symbol stream the decoder windows read (leftover of the previous push in
    // front), the lock search's persistent test buffer, per-window count of real tail symbols, and the DepuncXX / ViterbiSlidingBuffer
    // state of viterbi_punc.cpp / depunc.h carried between pushes
    bool punc = false;
    PuncTab ptab{};
    DevBuf<unsigned char> vitbuf, vitbuf_tmp, bdep;
    DevBuf<int> tail_real;
    DevBuf<PuncIdleOut> punc_out;
    PuncIdleOut *h_punc = nullptr;
    std::vector<int> h_tail, h_wcall;
    long in_buffer = 0;       // symbols waiting in vitbuf (ViterbiSlidingBuffer::in_buffer)
    int changing_shift = 0;   // DepuncXX::changing_shift
    bool is_first = false, got_extra = false;
    int buf_value = 128;      // DepuncXX::buf: the last symbol a call held back
    int test_bit_len = 0;     // Viterbi_Depunc::test_bit_len
    int pdec_start = -1;      // start state of the chained test decoder (cc_decoder_ber)
    int tb_blocks = 1;
    int tb_overlap = TB_OVERLAP; // warm-up rows of the parallel chainback blocks (adapts to the channel)
    int tb_clean = 0, spec_clean = 0; // consecutive launches without misses (the windows shrink back slowly)
    long tb_serial_total = 0, start_redone = 0; // chunks chained back serially / decoded again from a corrected start state
    DevBuf<VitIdleOut> idle_out;
    DevBuf<VitIdle2Out> idle2_out;
    VitIdle2Out *h_idle2 = nullptr;
    long idle_fallbacks = 0;
    DevBuf<DefrState> dstate;
    DevBuf<DefrEvent> devents;
    DevBuf<FrameRec> frames;
    DevBuf<uint8_t> frames_out, frames_tmp;
    DevBuf<RsTables> tables;
    // ccsds_simple_psk_decoder: second deframer (QPSK without NRZ-M runs one on the symbols as they are, one on the 90-degree
    // rotated ones) and the carried registers of the hard-decision stage
    DevBuf<uint32_t> chunk_bits2, fifo2;
    DevBuf<DefrState> dstate2;
    DevBuf<int> slice_carry; // [2][2]: oqpsk_delay register, QPSKDiff's previous symbol (ping-pong by call)
    DefrState h_dstate2{};
    long fifo_bits2 = 32;
    int defr_state2 = 2, slice_par = 0, last_nf = 0;
    void swap_unit();
    void compact_fifo();
    // pinned host mirrors
    VitRec *h_rec = nullptr;
    VitIdleOut *h_idle = nullptr;
    int *h_counters = nullptr;
    DefrState *h_dstate = nullptr;
    DefrEvent *h_events = nullptr;
    int *h_rs_err = nullptr;

    // stream state
    long soft_have = 0, max_chunks = 0, fifo_bits = 32, max_frames_push = 0;
    int vit_state = 0, invalid = 0, main_next_start = -1, enc_state = 0, nrzm_last = 0, nosync_runs = 0;
    VitHyp hyp{0, 0, 0};
    VitIdleState idle_st{-1, 0};
    float last_ber = 10.f;
    int defr_state_now = 2;
    long out_frames = 0; // frames waiting in frames_out
    long last_bits0 = 0, last_nbits = 0; // FIFO range of the bits appended by the last process() (debug)
    // statistics
    long total_soft = 0, total_chunks = 0, total_bits = 0, total_frames = 0, rs_corrected = 0, rs_failed = 0, replays = 0, launches = 0;

  private:
    long decode_chunks(long c0, long nch, bool single_step);
    void deframe_and_rs(long new_bits);
};

void build_rs_tables(RsTables &T);

} // namespace b200
