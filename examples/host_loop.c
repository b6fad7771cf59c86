/* A plain-C caller of the drop-in boundary (include/bgym.h): N envs on GPU 0 driven from host buffers with a
 * uniform random legal policy, no Python, no torch.
 *
 *   gcc -Iinclude -o host_loop examples/host_loop.c -Lbalatro_gym_b200 -lbgym -Wl,-rpath,$PWD/balatro_gym_b200
 *   ./host_loop 4096 2000
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

#include "bgym.h"

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint32_t next_u32(void) {            /* xorshift64*: the host policy's own randomness */
  rng_state ^= rng_state >> 12; rng_state ^= rng_state << 25; rng_state ^= rng_state >> 27;
  return (uint32_t)((rng_state * 0x2545F4914F6CDD1Dull) >> 32);
}

int main(int argc, char** argv) {
  const int64_t n = argc > 1 ? atoll(argv[1]) : 4096;
  const int steps = argc > 2 ? atoi(argv[2]) : 1000;
  if (bgym_device_count() < 1) { fprintf(stderr, "no CUDA device\n"); return 2; }
  BgymVec* v = NULL;
  if (bgym_vec_create(&v, n, 0)) { fprintf(stderr, "%s\n", bgym_last_error()); return 1; }
  uint32_t* seeds = malloc(n * sizeof *seeds);
  BgymObs* obs = malloc(n * sizeof *obs);
  int32_t* actions = malloc(n * sizeof *actions);
  double* reward = malloc(n * sizeof *reward);
  uint8_t* term = malloc(n), *trunc = malloc(n);
  for (int64_t i = 0; i < n; i++) seeds[i] = (uint32_t)(i + 1);
  if (bgym_vec_reset_host(v, seeds, NULL, obs)) { fprintf(stderr, "%s\n", bgym_last_error()); return 1; }
  double ret = 0.0; long episodes = 0;
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int s = 0; s < steps; s++) {
    for (int64_t i = 0; i < n; i++) {       /* k-th legal action of the packed mask word */
      uint64_t m = obs[i].action_mask_bits;
      int k = m ? (int)(((uint64_t)next_u32() * (uint32_t)__builtin_popcountll(m)) >> 32) : 0;
      while (k-- > 0) m &= m - 1;
      actions[i] = m ? __builtin_ctzll(m) : 0;
    }
    if (bgym_vec_step_host(v, actions, NULL, obs, reward, term, trunc, NULL, BGYM_FLAG_AUTORESET)) {
      fprintf(stderr, "%s\n", bgym_last_error()); return 1;
    }
    for (int64_t i = 0; i < n; i++) { ret += reward[i]; episodes += term[i]; }
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  const double sec = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
  printf("%lld envs x %d steps: %.3e env-steps/s, %ld episodes finished, mean reward per step %.4f\n",
         (long long)n, steps, (double)n * steps / sec, episodes, ret / ((double)n * steps));
  bgym_vec_destroy(v);
  free(seeds); free(obs); free(actions); free(reward); free(term); free(trunc);
  return 0;
}
