// state.cuh -- device-resident control block shared by the Arnoldi-step kernels.
//
// One Arnoldi expansion (decomposition.py:56-66 of the reference) is enqueued
// as a fixed sequence of kernels with no host round trip.  The data-dependent
// decisions of the reference -- the DGKS "repeat once" test (ortho.py:101) and
// the breakdown test (ortho.py:107, decomposition.py:61-63) -- are taken by the
// last block of the reducing kernel and published here; the kernels that follow
// read them and return at once when they have nothing to do.
#pragma once

#include "common.cuh"

namespace ab200 {

struct StepCtl {
  int stop;          // a step broke down (beta < tol): every later kernel is a no-op
  int broke_at;      // column j of that step, -1 otherwise
  int round2;        // the DGKS test of the current step asked for a second round
  int rounds_total;  // CGS/MGS rounds executed since reset
  int second_total;  // steps where the second round ran
  int steps_total;   // Arnoldi steps (operator applications) executed since reset
  int comm_error;    // multi-GPU wait timed out
  int pad;
  double nrm0sq;     // ||w||^2 entering the current round  (ortho.py:92 / :36)
  double beta;       // ||w|| after the latest update       (ortho.py:98,105 / :43,52)
};

// Peer-memory communicator: each rank owns `slots`, an array of
// kMaxRanks * slot_doubles doubles per sequence parity, that every peer can write
// over NVLink (CUDA IPC mapping).  A reduction is one remote store of the local
// partial into slot[my_rank] of every peer followed by a flag store; the reader
// sums the slots in rank order, so every rank gets the same bits.
constexpr int kMaxRanks = 8;
struct PeerComm {
  int nranks;
  int rank;
  int slot_doubles;                 // payload capacity of one slot
  int pad;
  double* slots[kMaxRanks];         // slots[r] = base of rank r's receive area (peer mapped)
  unsigned long long* flags[kMaxRanks];  // flags[r] = base of rank r's flag array
  unsigned long long* seq;          // local: monotonically increasing exchange counter
};

}  // namespace ab200
