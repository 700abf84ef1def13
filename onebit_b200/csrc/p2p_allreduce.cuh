// One-shot Lamport all-reduce over peer memory (p2p_allreduce.cu).
#pragma once
#include "common.cuh"

namespace onebit {

constexpr int kP2PMaxRanks = 8;

struct P2PComm {
    int rank, n;
    size_t cap;                   // floats per (rotation, rank) slot
    float* peer[kP2PMaxRanks];    // peer[r] = rank r's symmetric buffer [3][n][cap], addressable from this device
    unsigned* call_counter;       // device: number of collectives completed on this rank
    unsigned* cta_ticket;         // device: CTAs of the running collective that have finished
    unsigned* last_count;         // device [3]: elements the last call on each rotating region used (what must be re-armed)
    int* error_flag;              // device: 1 = a peer's data did not arrive within 2 s
};

int p2p_allreduce(const P2PComm& c, float* data, int64_t count, cudaStream_t s);

}  // namespace onebit
