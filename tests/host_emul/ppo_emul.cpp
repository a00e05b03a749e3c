// TEST INFRASTRUCTURE: the phase functions of the persistent PPO kernel (deeprl_b200/csrc/ppo_phases.h, ppo_sequence.inc)
// compiled for the host.  A "block" is emulated by running the NT threads of a phase one after another and the barrier by the
// end of that loop -- valid because no phase lets a thread read what another thread of the SAME phase writes (which is also what
// makes the CUDA version race-free).  tests/test_ppo_persistent.py builds this with g++ and checks it against the oracle's
// PPO update (oracle/agents.py ppo_update) without a GPU; the CUDA build of the same source is checked on the device.
#include <cstdint>
#include <vector>

#include "../../deeprl_b200/csrc/ppo_phases.h"

extern "C" int ppo_emul_minibatch_updates(const float* state, const float* action, const float* old_log_pi_a, const float* ret,
                                          const float* advantage, int32_t D, int32_t A, int32_t H1, int32_t H2, int32_t mb,
                                          const int64_t* perm, int32_t n_batches, float* a_flat, float* a_exp_avg,
                                          float* a_exp_avg_sq, int64_t* a_step, const int32_t* a_off, float* c_flat,
                                          float* c_exp_avg, float* c_exp_avg_sq, int64_t* c_step, const int32_t* c_off, float a_lr,
                                          float a_beta1, float a_beta2, float a_eps, float c_lr, float c_beta1, float c_beta2,
                                          float c_eps, float ratio_clip, float entropy_weight, float kl_gate, float* stats,
                                          int32_t n_threads) {
  using namespace b2rl_ppo;
  PpoArgs a;
  a.state = state; a.action = action; a.old_logp = old_log_pi_a; a.ret = ret; a.adv = advantage;
  a.D = D; a.A = A; a.H1 = H1; a.H2 = H2; a.mb = mb; a.perm = perm; a.n_batches = n_batches;
  a.a_flat = a_flat; a.a_m = a_exp_avg; a.a_v = a_exp_avg_sq; a.a_step = a_step;
  a.c_flat = c_flat; a.c_m = c_exp_avg; a.c_v = c_exp_avg_sq; a.c_step = c_step;
  for (int i = 0; i < 7; ++i) a.a_off[i] = a_off[i];
  for (int i = 0; i < 6; ++i) a.c_off[i] = c_off[i];
  a.a_lr = a_lr; a.a_b1 = a_beta1; a.a_b2 = a_beta2; a.a_eps = a_eps;
  a.c_lr = c_lr; a.c_b1 = c_beta1; a.c_b2 = c_beta2; a.c_eps = c_eps;
  a.clip = ratio_clip; a.ent_w = entropy_weight; a.gate_max = kl_gate; a.stats = stats; a.clk = nullptr;
  PpoShared S;
  float dummy[4];
  const size_t n = ppo_carve(S, dummy, D, A, H1, H2, mb);
  std::vector<float> block(n, -12345.0f);                 // (poisoned: a phase that reads before anybody wrote shows up)
  ppo_carve(S, block.data(), D, A, H1, H2, mb);
  const int NT = n_threads;
#define PPO_PHASE(stmt) for (int tid = 0; tid < NT; ++tid) { stmt; }
#include "../../deeprl_b200/csrc/ppo_sequence.inc"
#undef PPO_PHASE
  return 0;
}

// reverse thread order inside every phase: any dependence on intra-phase execution order changes the result
extern "C" int ppo_emul_minibatch_updates_reversed(const float* state, const float* action, const float* old_log_pi_a,
                                                   const float* ret, const float* advantage, int32_t D, int32_t A, int32_t H1,
                                                   int32_t H2, int32_t mb, const int64_t* perm, int32_t n_batches, float* a_flat,
                                                   float* a_exp_avg, float* a_exp_avg_sq, int64_t* a_step, const int32_t* a_off,
                                                   float* c_flat, float* c_exp_avg, float* c_exp_avg_sq, int64_t* c_step,
                                                   const int32_t* c_off, float a_lr, float a_beta1, float a_beta2, float a_eps,
                                                   float c_lr, float c_beta1, float c_beta2, float c_eps, float ratio_clip,
                                                   float entropy_weight, float kl_gate, float* stats, int32_t n_threads) {
  using namespace b2rl_ppo;
  PpoArgs a;
  a.state = state; a.action = action; a.old_logp = old_log_pi_a; a.ret = ret; a.adv = advantage;
  a.D = D; a.A = A; a.H1 = H1; a.H2 = H2; a.mb = mb; a.perm = perm; a.n_batches = n_batches;
  a.a_flat = a_flat; a.a_m = a_exp_avg; a.a_v = a_exp_avg_sq; a.a_step = a_step;
  a.c_flat = c_flat; a.c_m = c_exp_avg; a.c_v = c_exp_avg_sq; a.c_step = c_step;
  for (int i = 0; i < 7; ++i) a.a_off[i] = a_off[i];
  for (int i = 0; i < 6; ++i) a.c_off[i] = c_off[i];
  a.a_lr = a_lr; a.a_b1 = a_beta1; a.a_b2 = a_beta2; a.a_eps = a_eps;
  a.c_lr = c_lr; a.c_b1 = c_beta1; a.c_b2 = c_beta2; a.c_eps = c_eps;
  a.clip = ratio_clip; a.ent_w = entropy_weight; a.gate_max = kl_gate; a.stats = stats; a.clk = nullptr;
  PpoShared S;
  float dummy[4];
  const size_t n = ppo_carve(S, dummy, D, A, H1, H2, mb);
  std::vector<float> block(n, -12345.0f);
  ppo_carve(S, block.data(), D, A, H1, H2, mb);
  const int NT = n_threads;
#define PPO_PHASE(stmt) for (int tid = NT - 1; tid >= 0; --tid) { stmt; }
#include "../../deeprl_b200/csrc/ppo_sequence.inc"
#undef PPO_PHASE
  return 0;
}
