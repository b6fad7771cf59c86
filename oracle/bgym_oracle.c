/* bgym_oracle.c — CPU restatement of the reference's env step / hand scoring path.
 *
 * TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * link or call this file; the product (balatro_gym_b200/) never does.  It is a plain scalar C
 * restatement of cassiusfive/balatro-gym, one env at a time, written to read like the Python it
 * follows (lists, loops) and NOT like the CUDA kernels, so that the two are independent.
 *
 * Parity pinning: tests/test_oracle_vs_reference.py runs this file in lock-step against the
 * UNMODIFIED reference (oracle/refenv.py: BalatroEnv.step with every RNG tapped) and
 * tests/golden/ holds traces recorded from the reference itself
 * (tests/golden/make_golden.py).  The 8 surviving known-answer vectors of the reference's own
 * tests are in tests/test_known_answers.py.
 *
 * Each function cites the reference lines it restates (paths relative to the reference root).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../include/bgym.h"
#include "../include/bgym_tables.h"

static const uint8_t JOKER_COST[BGYM_NUM_JOKERS + 1] = BGYM_JOKER_COST_INIT;
static const int BASE_CHIPS[12] = BGYM_BASE_CHIPS_INIT;
static const int BASE_MULT[12] = BGYM_BASE_MULT_INIT;
static const int BLIND_CHIPS[8][3] = BGYM_BLIND_CHIPS_INIT;
static const int PACK_COST[5] = BGYM_PACK_COST_INIT;
static const int VOUCHER_COST[2] = BGYM_VOUCHER_COST_INIT;
static const double POW_1_15[101] = BGYM_POW_1_15_INIT;
static const double POW_0_8[9] = BGYM_POW_0_8_INIT;
static const double POW_1_5[101] = BGYM_POW_1_5_INIT;

/* ------------------------------------------------------------------------------------------
 * random draws: native Philox4x32-10 stream, or replay of the reference's recorded draws
 * ---------------------------------------------------------------------------------------- */
#define PHILOX_KEY1 0xB200CAFEu

static void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                          uint32_t out[4]) {
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

typedef struct Rng {
  BgymState* s;            /* native: seed/counter live in the state */
  const BgymDraws* tape;   /* replay: NULL in native mode */
  int iu, ik;
  uint32_t buf[4]; int pos; /* native: words of the current Philox block; pos == 4 -> empty.
                               Left-over words are dropped at the end of each step/reset call. */
} Rng;

static void rng_init(Rng* r, BgymState* s, const BgymDraws* tape) {
  r->s = s; r->tape = tape; r->iu = 0; r->ik = 0; r->pos = 4;
}

static uint32_t rng_word(Rng* r) {
  if (r->pos == 4) {
    philox4x32_10(r->s->rng_ctr++, 0, 0, 0, r->s->rng_seed, PHILOX_KEY1, r->buf);
    r->pos = 0;
  }
  return r->buf[r->pos++];
}

/* uniform in [0,1) with CPython's 53-bit construction (random.random: (a>>5, b>>6)) */
static double rng_u01(Rng* r) {
  if (r->tape) return r->tape->u[r->iu++];
  uint32_t a = rng_word(r) >> 5;
  uint32_t b = rng_word(r) >> 6;
  return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0);
}

/* unbiased uniform integer in [0,n): Lemire's multiply-shift with rejection on 32-bit words
 * (same distribution as CPython's _randbelow; the mapping from words differs, which is why parity
 * with the reference goes through replay) */
static int rng_below(Rng* r, int n) {
  if (r->tape) return r->tape->k[r->ik++];
  uint32_t un = (uint32_t)n;
  uint64_t m = (uint64_t)rng_word(r) * un;
  uint32_t l = (uint32_t)m;
  if (l < un) {
    uint32_t t = (0u - un) % un;
    while (l < t) {
      m = (uint64_t)rng_word(r) * un;
      l = (uint32_t)m;
    }
  }
  return (int)(m >> 32);
}

/* k distinct indices out of n, in draw order (random.sample semantics; replay gives them directly) */
static void rng_sample(Rng* r, int n, int k, int* out) {
  if (r->tape) {
    for (int i = 0; i < k; i++) out[i] = r->tape->k[r->ik++];
    return;
  }
  uint8_t taken[256];
  memset(taken, 0, sizeof taken);
  for (int t = 0; t < k; t++) {
    int j = rng_below(r, n - t);
    int idx = 0;
    for (;; idx++) {
      if (taken[idx]) continue;
      if (j == 0) break;
      j--;
    }
    taken[idx] = 1;
    out[t] = idx;
  }
}

/* ------------------------------------------------------------------------------------------
 * card helpers (cards.py)
 * ---------------------------------------------------------------------------------------- */
static int c16_code(uint16_t c) { return c & 63; }
static int c16_enh(uint16_t c) { return (c >> 6) & 15; }
static int c16_edition(uint16_t c) { return (c >> 10) & 7; }
static int c16_seal(uint16_t c) { return (c >> 13) & 7; }
static int code_rank(int code) { return code / 4 + 2; } /* cards.py:103 inverse */
static int code_suit(int code) { return code % 4; }

/* Rank.base_chips cards.py:52-60 */
static int rank_base_chips(int rank) {
  if (rank <= 10) return rank;
  if (rank == 14) return 11;
  return 10;
}

/* CardState.calculate_chip_bonus cards.py:262-267 via CardAdapter.to_scoring_format balatro_env_2.py:301 */
static int card_chip_value(int code, int enh, int edition) {
  int total = rank_base_chips(code_rank(code));
  if (enh == BGYM_ENH_BONUS) total += 30;
  else if (enh == BGYM_ENH_STONE) total += 50;
  if (edition == BGYM_ED_FOIL) total += 50;
  return total;
}

/* BalatroGame._classify_hand balatro_game.py:40-93 */
static int classify_hand(const int* codes, int n) {
  if (n == 0) return BGYM_HT_HIGH_CARD;
  int rank_counts[15] = {0}, suit_counts[4] = {0};
  for (int i = 0; i < n; i++) {
    rank_counts[code_rank(codes[i])]++;
    suit_counts[code_suit(codes[i])]++;
  }
  /* counts sorted descending: only the top two matter */
  int c0 = 0, c1 = 0, n_ranks = 0, n_suits = 0;
  for (int r = 2; r <= 14; r++) {
    int c = rank_counts[r];
    if (!c) continue;
    n_ranks++;
    if (c > c0) { c1 = c0; c0 = c; }
    else if (c > c1) c1 = c;
  }
  for (int s = 0; s < 4; s++) n_suits += suit_counts[s] > 0;
  int is_flush = (n_suits == 1) && n >= 5;
  int sorted_ranks[13], m = 0;
  for (int r = 2; r <= 14; r++) if (rank_counts[r]) sorted_ranks[m++] = r;
  int is_straight = 0;
  if (m >= 5) {
    for (int i = 0; i + 4 < m; i++)
      if (sorted_ranks[i + 4] - sorted_ranks[i] == 4) { is_straight = 1; break; }
    if (!is_straight && rank_counts[14] && rank_counts[2] && rank_counts[3] && rank_counts[4] && rank_counts[5])
      is_straight = 1;
  }
  if (is_straight && is_flush && n >= 5) return BGYM_HT_STRAIGHT_FLUSH;
  if (c0 == 4) return BGYM_HT_FOUR_KIND;
  if (n_ranks >= 2 && c0 == 3 && c1 == 2) return BGYM_HT_FULL_HOUSE;
  if (is_flush && n >= 5) return BGYM_HT_FLUSH;
  if (is_straight && n >= 5) return BGYM_HT_STRAIGHT;
  if (c0 == 3) return BGYM_HT_THREE_KIND;
  if (n_ranks >= 2 && c0 == 2 && c1 == 2) return BGYM_HT_TWO_PAIR;
  if (c0 == 2) return BGYM_HT_ONE_PAIR;
  return BGYM_HT_HIGH_CARD;
}

/* RULES evaluator: BalatroSimulator.evaluate_hand balatro_sim.py:220-400 with get_x_same :110-125,
 * get_flush :127-148, get_straight :150-214.  Returns the hand type named by results['top']. */
static int x_same_groups(const int* rank_counts, int num) { /* ranks held by EXACTLY num cards (:118-119) */
  int groups = 0;
  for (int r = 2; r <= 14; r++) groups += rank_counts[r] == num;
  return groups;
}

static int rules_flush(const int* codes, int n, int four_fingers) {
  int required = four_fingers ? 4 : 5;
  if (n > 5 || n < required) return 0;                       /* :133-134 */
  for (int suit = 0; suit < 4; suit++) {
    int count = 0;
    for (int i = 0; i < n; i++) count += code_suit(codes[i]) == suit;
    if (count >= required) return 1;
  }
  return 0;
}

static int rules_straight(const int* rank_counts, int n, int four_fingers, int shortcut) {
  int required = four_fingers ? 4 : 5;
  if (n > 5 || n < required) return 0;                       /* :155-156 */
  int length = 0, skipped = 0;
  for (int r = 14; r > 1; r--) {                             /* :173-187 */
    if (rank_counts[r]) length++;
    else if (shortcut && !skipped) skipped = 1;
    else { length = 0; skipped = 0; }
    if (length >= required) return 1;
  }
  static const int WHEEL[5] = {14, 2, 3, 4, 5};              /* :190-206; `skipped` carries over from the scan */
  int wheel = 0;
  for (int k = 0; k < 5; k++) {
    if (rank_counts[WHEEL[k]]) wheel++;
    else if (shortcut && !skipped) skipped = 1;
    else break;
  }
  return wheel >= required;
}

static int classify_rules(const int* codes, int n, int four_fingers, int shortcut) {
  int rank_counts[15] = {0};
  for (int i = 0; i < n; i++) rank_counts[code_rank(codes[i])]++;
  int n5 = x_same_groups(rank_counts, 5), n4 = x_same_groups(rank_counts, 4);
  int n3 = x_same_groups(rank_counts, 3), n2 = x_same_groups(rank_counts, 2);
  int flush = rules_flush(codes, n, four_fingers);
  int straight = rules_straight(rank_counts, n, four_fingers, shortcut);
  if (n5 && flush) return BGYM_HT_FLUSH_FIVE;               /* :255-258, in priority order down to :333 */
  if (n3 && n2 && flush) return BGYM_HT_FLUSH_HOUSE;
  if (n5) return BGYM_HT_FIVE_KIND;
  if (flush && straight) return BGYM_HT_STRAIGHT_FLUSH;
  if (n4) return BGYM_HT_FOUR_KIND;
  if (n3 && n2) return BGYM_HT_FULL_HOUSE;
  if (flush) return BGYM_HT_FLUSH;
  if (straight) return BGYM_HT_STRAIGHT;
  if (n3) return BGYM_HT_THREE_KIND;
  if (n2 == 2 || (n3 == 1 && n2 == 1)) return BGYM_HT_TWO_PAIR;   /* exactly two pairs: three pairs fall through */
  if (n2) return BGYM_HT_ONE_PAIR;
  return BGYM_HT_HIGH_CARD;
}

/* ScoreEngine.get_hand_chips_mult scoring_engine.py:87-101 (engine level is capped at 15,
 * apply_planet :82-85; state.hand_levels is not, balatro_env_2.py:1119) */
static void hand_chips_mult(const uint8_t* level, int ht, int* chips, int* mult) {
  int lv = level[ht];
  if (lv > 15) lv = 15;
  if (lv < 1) lv = 1;
  *chips = BASE_CHIPS[ht] + (lv - 1) * 10;
  *mult = BASE_MULT[ht] + (lv - 1);
}

/* ------------------------------------------------------------------------------------------
 * joker pipeline: UnifiedScorer.score_hand unified_scoring.py:111-299 with
 * CompleteJokerEffects complete_joker_effects.py:35-183 written out joker by joker
 * ---------------------------------------------------------------------------------------- */
typedef struct ScoreCard { int rank; int suit; /* suit -1 = 'Stone' */ int chip_value; } ScoreCard;

typedef struct ScoreIn {
  const ScoreCard* cards; int n_cards;
  const uint8_t* jokers; int n_jokers;       /* joker ids in order; unknown ids are inert */
  int hand_type; int table_names;            /* naming convention passed as context['hand_type'] */
  int hands_left, discards_left, deck_len;
  /* draws */
  const BgymScoreCtx* replay;                /* non-NULL: use recorded draws */
  uint32_t seed; uint64_t index;             /* native: Philox keyed by (seed, hand index) */
  uint32_t ctr; uint32_t buf[4]; int pos;
} ScoreIn;

typedef struct ScoreOut { int chips, mult; double x_mult; int64_t score; int money; } ScoreOut;

/* which joker-table hand name does the context name equal?  The env passes
 * hand_type.name.replace('_',' ').title() (balatro_env_2.py:674): 'One Pair', 'Three Kind', 'Four Kind'
 * never equal the table's 'Pair', 'Three of a Kind', 'Four of a Kind' (SURVEY Q13). */
static int name_matches(int ht, int table_names, int hn) {
  switch (hn) {
    case BGYM_HN_PAIR: return table_names && ht == BGYM_HT_ONE_PAIR;
    case BGYM_HN_THREE_OAK: return table_names && ht == BGYM_HT_THREE_KIND;
    case BGYM_HN_FOUR_OAK: return table_names && ht == BGYM_HT_FOUR_KIND;
    case BGYM_HN_TWO_PAIR: return ht == BGYM_HT_TWO_PAIR;
    case BGYM_HN_STRAIGHT: return ht == BGYM_HT_STRAIGHT;
    case BGYM_HN_FLUSH: return ht == BGYM_HT_FLUSH;
  }
  return 0;
}

static uint32_t score_word(ScoreIn* in) {
  if (in->pos == 4) {
    philox4x32_10(in->ctr++, (uint32_t)in->index, (uint32_t)(in->index >> 32), 1, in->seed, PHILOX_KEY1, in->buf);
    in->pos = 0;
  }
  return in->buf[in->pos++];
}
static double score_u01(ScoreIn* in) {
  uint32_t a = score_word(in) >> 5;
  uint32_t b = score_word(in) >> 6;
  return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0);
}
static int score_below(ScoreIn* in, int n) {
  uint32_t un = (uint32_t)n;
  uint64_t m = (uint64_t)score_word(in) * un;
  uint32_t l = (uint32_t)m;
  if (l < un) {
    uint32_t t = (0u - un) % un;
    while (l < t) { m = (uint64_t)score_word(in) * un; l = (uint32_t)m; }
  }
  return (int)(m >> 32);
}

static void score_hand(ScoreIn* in, ScoreOut* out, const uint8_t* levels) {
  int chips, mult;
  hand_chips_mult(levels, in->hand_type, &chips, &mult);          /* :120 */
  double x_mult = 1.0;
  int money = 0;
  for (int i = 0; i < in->n_cards; i++) chips += in->cards[i].chip_value; /* :139-153 */

  /* individual phase :173-209 — card-major, joker-minor; chips/mult are sums, x_mult a running product */
  int ind_chips = 0, ind_mult = 0;
  double ind_x = 1.0;
  int misprint_seen = 0;
  for (int c = 0; c < in->n_cards; c++) {
    int rank = in->cards[c].rank, suit = in->cards[c].suit;
    int is_face = rank == 11 || rank == 12 || rank == 13;
    for (int j = 0; j < in->n_jokers; j++) {
      int ec = 0, em = 0, emoney = 0;
      double ex = 1.0;
      switch (in->jokers[j]) {
        case BGYM_J_FIBONACCI: if (rank == 2 || rank == 3 || rank == 5 || rank == 8 || rank == 14) em = 8; break;
        case BGYM_J_EVEN_STEVEN: if (rank == 2 || rank == 4 || rank == 6 || rank == 8 || rank == 10) em = 4; break;
        case BGYM_J_ODD_TODD: if (rank == 3 || rank == 5 || rank == 7 || rank == 9 || rank == 14) ec = 31; break;
        case BGYM_J_SCHOLAR: if (rank == 14) { ec = 20; em = 4; } break;
        case BGYM_J_WALKIE_TALKIE: if (rank == 4 || rank == 10) { ec = 10; em = 4; } break;
        case BGYM_J_WEE_JOKER: if (rank == 2) ec = 8; break;
        case BGYM_J_SCARY_FACE: if (is_face) ec = 30; break;
        case BGYM_J_SMILEY_FACE: if (is_face) em = 5; break;
        case BGYM_J_TRIBOULET: if (rank == 12 || rank == 13) ex = 2.0; break;
        case BGYM_J_ARROWHEAD: if (suit == 3) ec = 50; break;
        case BGYM_J_ONYX_AGATE: if (suit == 0) em = 7; break;
        case BGYM_J_ROUGH_GEM: if (suit == 1) emoney = 1; break;
        case BGYM_J_BLOODSTONE: {
          /* one roll per (card, Bloodstone) pair, used only for Hearts (:161, :179-182) */
          int hit;
          if (in->replay) hit = (in->replay->bloodstone_bits >> c) & 1;
          else hit = score_u01(in) < 0.5;
          if (suit == 2 && hit) ex = 2.0;
          break;
        }
        default: break; /* 8 Ball: no numeric effect (:147, :167-170) */
      }
      ind_chips += ec; ind_mult += em; ind_x *= ex; money += emoney;
    }
  }
  chips += ind_chips; mult += ind_mult; x_mult *= ind_x;

  /* main phase :211-244, joker order */
  int suit_present[4] = {0, 0, 0, 0}, n_suits_present = 0, kings = 0, queens = 0, stone_present = 0;
  for (int c = 0; c < in->n_cards; c++) {
    int s = in->cards[c].suit;
    if (s >= 0) suit_present[s] = 1; else stone_present = 1;
    if (in->cards[c].rank == 13) kings++;
    if (in->cards[c].rank == 12) queens++;
  }
  for (int s = 0; s < 4; s++) n_suits_present += suit_present[s];
  int n_distinct_suit_names = n_suits_present + stone_present; /* 'Stone' is a suit string too */
  for (int j = 0; j < in->n_jokers; j++) {
    int ec = 0, em = 0;
    double ex = 1.0;
    switch (in->jokers[j]) {
      case BGYM_J_JOKER: em = 4; break;
      case BGYM_J_STUNTMAN: ec = 250; break;
      case BGYM_J_MISPRINT:
        if (in->replay) em = in->replay->misprint[misprint_seen < 5 ? misprint_seen : 4];
        else em = score_below(in, 24);
        misprint_seen++;
        break;
      case BGYM_J_GROS_MICHEL: em = 15; break;
      case BGYM_J_CAVENDISH: ex = 3.0; break;
      case BGYM_J_HALF_JOKER: if (in->n_cards <= 3) em = 20; break;
      case BGYM_J_ABSTRACT_JOKER: em = 3 * in->n_jokers; break;
      case BGYM_J_ACROBAT: if (in->hands_left == 1) ex = 3.0; break;
      case BGYM_J_MYSTIC_SUMMIT: if (in->discards_left == 0) em = 15; break;
      case BGYM_J_BANNER: ec = 30 * in->discards_left; break;
      case BGYM_J_BLUE_JOKER: ec = 2 * in->deck_len; break;
      case BGYM_J_POPCORN: em = 20; break;
      case BGYM_J_ICE_CREAM: ec = 100; break;
      case BGYM_J_GREEDY_JOKER: if (suit_present[1]) em = 3; break;
      case BGYM_J_LUSTY_JOKER: if (suit_present[2]) em = 3; break;
      case BGYM_J_WRATHFUL_JOKER: if (suit_present[3]) em = 3; break;
      case BGYM_J_GLUTTONOUS_JOKER: if (suit_present[0]) em = 3; break;
      case BGYM_J_JOLLY_JOKER: if (name_matches(in->hand_type, in->table_names, BGYM_HN_PAIR)) em = 8; break;
      case BGYM_J_ZANY_JOKER: if (name_matches(in->hand_type, in->table_names, BGYM_HN_THREE_OAK)) em = 12; break;
      case BGYM_J_MAD_JOKER: if (name_matches(in->hand_type, in->table_names, BGYM_HN_TWO_PAIR)) em = 10; break;
      case BGYM_J_CRAZY_JOKER: if (name_matches(in->hand_type, in->table_names, BGYM_HN_STRAIGHT)) em = 12; break;
      case BGYM_J_DROLL_JOKER: if (name_matches(in->hand_type, in->table_names, BGYM_HN_FLUSH)) em = 10; break;
      case BGYM_J_SLY_JOKER: if (name_matches(in->hand_type, in->table_names, BGYM_HN_PAIR)) ec = 50; break;
      case BGYM_J_WILY_JOKER: if (name_matches(in->hand_type, in->table_names, BGYM_HN_THREE_OAK)) ec = 100; break;
      case BGYM_J_CLEVER_JOKER: if (name_matches(in->hand_type, in->table_names, BGYM_HN_TWO_PAIR)) ec = 80; break;
      case BGYM_J_DEVIOUS_JOKER: if (name_matches(in->hand_type, in->table_names, BGYM_HN_STRAIGHT)) ec = 100; break;
      case BGYM_J_CRAFTY_JOKER: if (name_matches(in->hand_type, in->table_names, BGYM_HN_FLUSH)) ec = 80; break;
      case BGYM_J_THE_DUO: if (name_matches(in->hand_type, in->table_names, BGYM_HN_PAIR)) ex = 2.0; break;
      case BGYM_J_THE_TRIO: if (name_matches(in->hand_type, in->table_names, BGYM_HN_THREE_OAK)) ex = 3.0; break;
      case BGYM_J_THE_FAMILY: if (name_matches(in->hand_type, in->table_names, BGYM_HN_FOUR_OAK)) ex = 4.0; break;
      case BGYM_J_THE_ORDER: if (name_matches(in->hand_type, in->table_names, BGYM_HN_STRAIGHT)) ex = 3.0; break;
      case BGYM_J_THE_TRIBE: if (name_matches(in->hand_type, in->table_names, BGYM_HN_FLUSH)) ex = 2.0; break;
      case BGYM_J_BLACKBOARD: { /* all(card.suit in ['Spades','Clubs']) over context['cards'] :99-103 */
        int ok = 1;
        for (int c = 0; c < in->n_cards; c++) {
          int s = in->cards[c].suit;
          if (!(s == 3 || s == 0)) ok = 0;
        }
        if (ok) ex = 3.0;
        break;
      }
      case BGYM_J_SEEING_DOUBLE: /* 'Clubs' in suits and len(suits) > 1 :105-109 */
        if (suit_present[0] && n_distinct_suit_names > 1) ex = 2.0;
        break;
      case BGYM_J_FLOWER_POT: /* len(suits) == 4 — 'Stone' counts as a suit string :111-115 */
        if (n_distinct_suit_names == 4) ex = 3.0;
        break;
      case BGYM_J_BARON: if (kings > 0) ex = POW_1_5[kings]; break;         /* :117-121 */
      case BGYM_J_SHOOT_THE_MOON: if (queens > 0) em = 13 * queens; break;  /* :123-127 */
      default: break;
    }
    chips += ec; mult += em; x_mult *= ex;
  }
  out->chips = chips; out->mult = mult; out->x_mult = x_mult; out->money = money;
  /* final_score = int(chips * mult * x_mult) :286 — exact integer product, one fp64 multiply, truncate */
  out->score = (int64_t)((double)((int64_t)chips * (int64_t)mult) * x_mult);
}

int oracle_score_hands(const uint8_t* cards8, const uint16_t* mods8, const uint8_t* n_cards,
                       const uint8_t* jokers8, const uint8_t* levels12, const BgymScoreCtx* ctx,
                       uint8_t* hand_type, int32_t* chips, int32_t* mult, double* x_mult,
                       int64_t* score, int32_t* money, uint32_t seed, int64_t n, int flags) {
  static const uint8_t ones[12] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1};
  for (int64_t i = 0; i < n; i++) {
    int nc = n_cards ? n_cards[i] : 5;
    int codes[8];
    ScoreCard sc[8];
    for (int c = 0; c < nc; c++) {
      int code = cards8[i * 8 + c];
      int m = mods8 ? mods8[i * 8 + c] : 0;
      int enh = m & 15, ed = (m >> 4) & 15;
      codes[c] = code;
      sc[c].chip_value = card_chip_value(code, enh, ed);
      if (enh == BGYM_ENH_STONE) { sc[c].rank = 0; sc[c].suit = -1; }   /* balatro_env_2.py:304-306 */
      else { sc[c].rank = code_rank(code); sc[c].suit = code_suit(code); }
    }
    uint8_t jk[8];
    int nj = 0;
    if (jokers8) for (int j = 0; j < 8; j++) if (jokers8[i * 8 + j]) jk[nj++] = jokers8[i * 8 + j];
    ScoreIn in;
    memset(&in, 0, sizeof in);
    in.cards = sc; in.n_cards = nc; in.jokers = jk; in.n_jokers = nj;
    if (flags & BGYM_SCORE_RULES) {
      int four_fingers = 0, shortcut = 0;
      for (int j = 0; j < nj; j++) { four_fingers |= jk[j] == BGYM_J_FOUR_FINGERS; shortcut |= jk[j] == BGYM_J_SHORTCUT; }
      in.hand_type = classify_rules(codes, nc, four_fingers, shortcut);
    } else {
      in.hand_type = classify_hand(codes, nc);
    }
    in.table_names = (flags & BGYM_SCORE_TABLE_NAMES) != 0;
    in.hands_left = ctx ? ctx[i].hands_left : 4;
    in.discards_left = ctx ? ctx[i].discards_left : 3;
    in.deck_len = ctx ? ctx[i].deck_len : 52;
    in.replay = (ctx && ctx[i].use_replay) ? &ctx[i] : NULL;
    in.seed = seed; in.index = (uint64_t)i; in.ctr = 0; in.pos = 4;
    ScoreOut out;
    score_hand(&in, &out, levels12 ? levels12 + i * 12 : ones);
    hand_type[i] = (uint8_t)in.hand_type;
    chips[i] = out.chips; mult[i] = out.mult;
    if (x_mult) x_mult[i] = out.x_mult;
    score[i] = out.score;
    if (money) money[i] = out.money;
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * env: list helpers over the hand
 * ---------------------------------------------------------------------------------------- */
/* hand_code[i] = card code shown by obs['hand'][i] (cache kept in the hot record), 0xFF = -1 */
static void refresh_hand_codes(BgymState* s) {
  for (int i = 0; i < 8; i++) {
    if (i >= s->hand_n) s->hand[i] = 0xFF;
    if (i < s->hand_n && s->hand[i] < s->deck_n) s->hand_code[i] = (uint8_t)c16_code(s->deck[s->hand[i]]);
    else s->hand_code[i] = 0xFF;
  }
}

static int hand_contains(const BgymState* s, int deck_idx) {
  for (int i = 0; i < s->hand_n; i++) if (s->hand[i] == deck_idx) return 1;
  return 0;
}

static void hand_pop(BgymState* s, int slot) {
  for (int i = slot; i + 1 < s->hand_n; i++) s->hand[i] = s->hand[i + 1];
  s->hand_n--;
  s->hand[s->hand_n] = 0xFF;
}

/* BalatroGame._draw_cards balatro_game.py:95-109: top up with the LOWEST deck indices not in hand */
static void draw_cards(BgymState* s) {
  int want = (int)s->hand_size - (int)s->hand_n;
  for (int i = 0; i < s->deck_n && want > 0 && s->hand_n < 8; i++) {
    if (hand_contains(s, i)) continue;
    s->hand[s->hand_n++] = (uint8_t)i;
    want--;
  }
}

static int sel_slot(const BgymState* s, int k) { return (s->sel_order >> (4 * k)) & 15; }

/* ------------------------------------------------------------------------------------------
 * action mask  balatro_env_2.py:1426-1471
 * ---------------------------------------------------------------------------------------- */
static uint64_t action_mask(const BgymState* s) {
  uint64_t m = 0;
  if (s->phase == BGYM_PHASE_PLAY) {
    int n = s->hand_n < 8 ? s->hand_n : 8;
    for (int i = 0; i < n; i++) m |= 1ull << (BGYM_A_SELECT_BASE + i);
    if (s->sel_n > 0) m |= 1ull << BGYM_A_PLAY_HAND;
    if (s->sel_n > 0 && s->discards_left > 0) m |= 1ull << BGYM_A_DISCARD;
    for (int i = 0; i < s->cons_n; i++) m |= 1ull << (BGYM_A_USE_CONS_BASE + i);
  } else if (s->phase == BGYM_PHASE_SHOP) {
    /* `if self.shop:` — a Shop exists whenever the phase is SHOP */
    for (int i = 0; i < s->n_items; i++)
      if (s->money >= s->item_cost[i]) m |= 1ull << (BGYM_A_SHOP_BUY_BASE + i);
    if (s->money >= s->shop_reroll_state) m |= 1ull << BGYM_A_SHOP_REROLL;
    m |= 1ull << BGYM_A_SHOP_END;
    for (int i = 0; i < s->joker_n; i++) m |= 1ull << (BGYM_A_SELL_JOKER_BASE + i);
  } else if (s->phase == BGYM_PHASE_BLIND_SELECT) {
    for (int i = 0; i < 3; i++) m |= 1ull << (BGYM_A_SELECT_BLIND_BASE + i);
    m |= 1ull << BGYM_A_SKIP_BLIND;
  }
  return m;
}

int oracle_action_mask(const BgymState* state, uint64_t* mask, int64_t n) {
  for (int64_t i = 0; i < n; i++) mask[i] = action_mask(&state[i]);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * observation  balatro_env_2.py:1473-1573
 * ---------------------------------------------------------------------------------------- */
/* out-of-range Python ints: the reference's pinned numpy 1.26.4 wraps (two's complement, with a
 * DeprecationWarning) where numpy 2.x raises OverflowError; we follow the pinned behaviour. */
static int16_t wrap_i16(int64_t v) { return (int16_t)(uint16_t)(uint64_t)v; }
static int32_t clamp_i32(int64_t v) { return (int32_t)(v > 2147483647LL ? 2147483647LL : v); }

static int obs_consumable_id(int cid) { return cid >= BGYM_CONS_ENUMSTYLE_BASE ? 0 : cid; } /* :1571 */

static void write_obs(const BgymState* s, BgymObs* o) {
  memset(o, 0, sizeof *o);
  for (int i = 0; i < 8; i++) {
    o->hand[i] = -1;
    if (i < s->hand_n && s->hand[i] < s->deck_n) o->hand[i] = (int8_t)c16_code(s->deck[s->hand[i]]);
  }
  for (int k = 0; k < s->sel_n; k++) { int sl = sel_slot(s, k); if (sl < 8) o->selected_cards[sl] = 1; }
  for (int i = 0; i < 8; i++) o->face_down_cards[i] = (s->face_down_mask >> i) & 1;
  o->chips_scored = s->chips_scored;
  o->round_chips_scored = (int32_t)(uint32_t)(uint64_t)s->round_chips;
  {
    double needed = (double)(s->chips_needed > 1 ? s->chips_needed : 1);
    double p = (double)s->round_chips / needed;
    o->progress_ratio = (float)(p < 2.0 ? p : 2.0);
  }
  o->mult = 1;
  o->chips_needed = s->chips_needed;
  o->money = s->money;
  o->hands_played = s->hands_played_total;
  o->best_hand_this_ante = s->best_hand;
  o->ante = s->ante;
  o->shop_rerolls = wrap_i16(s->shop_reroll_state);
  for (int i = 0; i < s->joker_n && i < 10; i++) o->joker_ids[i] = s->joker_id[i];
  for (int i = 0; i < s->cons_n && i < 5; i++) o->consumables[i] = (int16_t)obs_consumable_id(s->cons_id[i]);
  if (s->phase == BGYM_PHASE_SHOP) {
    for (int i = 0; i < s->n_items; i++) {
      o->shop_items[i] = s->item_type[i];
      o->shop_costs[i] = wrap_i16(s->item_cost[i]);
    }
  }
  for (int i = 0; i < 12; i++) o->hand_levels[i] = (int8_t)s->hand_level[i];
  o->hand_size = (int8_t)s->hand_n;
  o->deck_size = (int8_t)s->deck_n;
  o->round = (int8_t)s->round;
  o->hands_left = (int8_t)s->hands_left;
  o->discards_left = (int8_t)s->discards_left;
  o->joker_count = (int8_t)s->joker_n;
  o->joker_slots = (int8_t)s->joker_slots;
  o->consumable_count = (int8_t)s->cons_n;
  o->consumable_slots = (int8_t)s->cons_slots;
  o->phase = (int8_t)s->phase;
  o->boss_blind_active = s->boss_type != 0;
  o->boss_blind_type = (int8_t)s->boss_type;
  uint64_t m = action_mask(s);
  o->action_mask_bits = m;
  /* the int8[60] action_mask of the reference (:1522) is carried as action_mask_bits only */
}

/* ------------------------------------------------------------------------------------------
 * reset  balatro_env_2.py:505-558, UnifiedGameState defaults :166-211
 * ---------------------------------------------------------------------------------------- */
static uint32_t next_episode_seed(uint32_t seed) { /* autoreset: seed of the following episode */
  uint32_t x = seed + 0x9E3779B9u;
  x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
  return x ? x : 1u;
}

/* Synthetic-state generator of BASELINE configs[2]/[3] (BGYM_FLAG_GEN_C3 / BGYM_FLAG_GEN_CONS).  This is
 * not reference code: the reference has no generator (SURVEY 8(d) C3/C4 defines the distributions, Appendix E
 * the injection recipe); the law below is the one include/bgym.h specifies, restated with lists. */
static uint32_t gen_scaled(uint32_t word, uint32_t n) { return (uint32_t)(((uint64_t)word * n) >> 32); }

static uint16_t gen_card_mods(uint32_t seed, int k) { /* card k = suit * 13 + rank - 2; two cards share a block */
  uint32_t w[4];
  philox4x32_10((uint32_t)(k / 2), 0, 0, 0, seed, BGYM_GEN_KEY1, w);
  uint32_t first = w[(k % 2) * 2], second = w[(k % 2) * 2 + 1];
  unsigned enh = 0, ed = 0, seal = 0;
  if ((first >> 30) == 0) enh = 1 + ((first >> 27) & 7);         /* 1/4, then uniform over the 8 enhancements */
  uint32_t e = (uint32_t)(((uint64_t)(first & 0x07FFFFFFu) * 30) >> 27);
  if (e < 3) ed = 1 + e;                                         /* 1/10 over FOIL, HOLO, POLY (low 27 bits) */
  uint32_t t = gen_scaled(second, 40); if (t < 4) seal = 1 + t;  /* 1/10 over the four seals */
  return (uint16_t)((enh << 6) | (ed << 10) | (seal << 13));
}

static void gen_jokers_consumables(BgymState* s, uint32_t seed, int flags) {
  uint32_t a[4], b[4];
  philox4x32_10(64, 0, 0, 0, seed, BGYM_GEN_KEY1, a);
  philox4x32_10(65, 0, 0, 0, seed, BGYM_GEN_KEY1, b);
  uint32_t words[5] = {a[0], a[1], a[2], a[3], b[0]};
  int pool[BGYM_NUM_SHOP_JOKERS], n_pool = BGYM_NUM_SHOP_JOKERS;   /* ids with base_cost > 0, ascending */
  for (int i = 0; i < n_pool; i++) pool[i] = i + 1;
  for (int t = 0; t < 5; t++) {                                    /* sample without replacement, draw order */
    int p = (int)gen_scaled(words[t], (uint32_t)n_pool);
    s->joker_id[t] = (uint8_t)pool[p];
    for (int i = p; i + 1 < n_pool; i++) pool[i] = pool[i + 1];
    n_pool--;
  }
  s->joker_n = 5;
  if (flags & BGYM_FLAG_GEN_CONS) {
    static const uint8_t ALL_IDS[52] = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22,
                                        30, 31, 32, 33, 34, 35, 36, 37, 38, 39, 40, 41,
                                        50, 51, 52, 53, 54, 55, 56, 57, 58, 59, 60, 61, 62, 63, 64, 65, 66, 67};
    s->cons_id[0] = ALL_IDS[gen_scaled(b[1], 52)];
    s->cons_id[1] = ALL_IDS[gen_scaled(b[2], 52)];
    s->cons_n = 2;
  }
}

static void reset_env(BgymState* s, uint32_t seed, const uint8_t* deck52, int flags) {
  memset(s, 0, sizeof *s);
  s->ante = 1; s->round = 1; s->phase = BGYM_PHASE_BLIND_SELECT;
  s->chips_needed = 300; s->money = 4;
  s->hands_left = 4; s->discards_left = 3; s->hand_size = 8;
  s->joker_slots = 5; s->cons_slots = 2;
  s->shop_reroll_state = 5;
  for (int i = 0; i < 12; i++) s->hand_level[i] = 1;
  s->deck_n = 52;
  memset(s->hand, 0xFF, 8);
  memset(s->hand_code, 0xFF, 8);
  s->rng_seed = seed; s->rng_ctr = 0;
  const int gen = (flags & BGYM_FLAG_GEN_C3) != 0;
  if (gen) gen_jokers_consumables(s, seed, flags);
  if (deck52) {
    for (int i = 0; i < 52; i++) {
      int rank = code_rank(deck52[i]), suit = code_suit(deck52[i]);
      s->deck[i] = (uint16_t)(deck52[i] | (gen ? gen_card_mods(seed, suit * 13 + rank - 2) : 0));
    }
  } else {
    /* suit-major, rank-minor build (:519-522) then Fisher-Yates exactly as random.shuffle:
     * for i in reversed(range(1, n)): j = randbelow(i + 1); swap */
    int k = 0;
    for (int suit = 0; suit < 4; suit++)
      for (int rank = 2; rank <= 14; rank++, k++)
        s->deck[k] = (uint16_t)(((rank - 2) * 4 + suit) | (gen ? gen_card_mods(seed, k) : 0));
    /* native draws of the shuffle: j_i comes from Philox block (i-1)/2 keyed (seed, SHUFFLE key),
     * words (0,1) for odd i, (2,3) for even i, so the 51 draws are independent of each other (the
     * kernel computes them lane-parallel); bounded by Lemire's multiply-shift, second word on the
     * (probability < n/2^32) rejection.  The step stream (rng_ctr) is not touched. */
    for (int i = 51; i >= 1; i--) {
      uint32_t w[4];
      philox4x32_10((uint32_t)((i - 1) / 2), 0, 0, 0, seed, 0xB200DECCu, w);
      uint32_t w0 = w[((i - 1) & 1) * 2], w1 = w[((i - 1) & 1) * 2 + 1];
      uint32_t un = (uint32_t)(i + 1);
      uint64_t m = (uint64_t)w0 * un;
      if ((uint32_t)m < (0u - un) % un) m = (uint64_t)w1 * un;
      int j = (int)(m >> 32);
      uint16_t t = s->deck[i]; s->deck[i] = s->deck[j]; s->deck[j] = t;
    }
  }
}

int oracle_reset(BgymState* state, BgymObs* obs, const uint8_t* reset_mask, const uint32_t* seeds,
                 const uint8_t* decks52, int64_t n, int flags) {
  for (int64_t i = 0; i < n; i++) {
    if (reset_mask && !reset_mask[i]) continue;
    reset_env(&state[i], seeds[i], decks52 ? decks52 + i * 52 : NULL, flags);
    if (obs && !(flags & BGYM_FLAG_NO_OBS)) write_obs(&state[i], &obs[i]);
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * shop  shop.py:96-205, balatro_env_2.py:1383-1392
 * ---------------------------------------------------------------------------------------- */
static double shop_cost_mult(const BgymState* s) { /* shop.py:105-109 */
  double m = POW_1_15[s->ante - 1];
  if (s->n_magic_trick > 0) m *= 0.9;
  return m;
}

static int joker_owned(const BgymState* s, int id) {
  for (int i = 0; i < s->joker_n; i++) if (s->joker_id[i] == id) return 1;
  return 0;
}

static void shop_generate_inventory(BgymState* s, Rng* r) { /* shop.py:112-139 */
  double mult = shop_cost_mult(s);
  int n = 0;
  int third = BGYM_PACK_TAROT + rng_below(r, 3);
  int packs[3] = {BGYM_PACK_STANDARD, BGYM_PACK_JOKER, third};
  for (int i = 0; i < 3; i++) {
    s->item_type[n] = BGYM_ITEM_PACK; s->item_id[n] = (uint8_t)packs[i];
    s->item_cost[n] = (int32_t)(PACK_COST[packs[i]] * mult);
    n++;
  }
  int candid[BGYM_NUM_JOKERS], nc = 0;
  for (int id = 1; id <= BGYM_NUM_JOKERS; id++)
    if (JOKER_COST[id] > 0 && !joker_owned(s, id)) candid[nc++] = id;
  int k = nc < 3 ? nc : 3, pick[3];
  rng_sample(r, nc, k, pick);
  for (int i = 0; i < k; i++) {
    int id = candid[pick[i]];
    s->item_type[n] = BGYM_ITEM_JOKER; s->item_id[n] = (uint8_t)id;
    s->item_cost[n] = (int32_t)(JOKER_COST[id] * mult);
    n++;
  }
  int v = rng_below(r, 2);
  s->item_type[n] = BGYM_ITEM_VOUCHER; s->item_id[n] = (uint8_t)v;
  s->item_cost[n] = (int32_t)(VOUCHER_COST[v] * mult);
  n++;
  for (int i = 0; i < 2; i++) {
    int c = rng_below(r, 52);
    s->item_type[n] = BGYM_ITEM_CARD; s->item_id[n] = (uint8_t)c; s->item_cost[n] = BGYM_CARD_COST;
    n++;
  }
  for (int i = n; i < 9; i++) { s->item_type[i] = 0; s->item_id[i] = 0; s->item_cost[i] = 0; }
  s->n_items = (uint8_t)n;
}

static void generate_shop(BgymState* s, Rng* r) { /* balatro_env_2.py:1383-1392 */
  /* the shop seed (rng stream 2) only seeds the Shop's own generator: no draw is consumed here */
  s->reroll_cost = BGYM_REROLL_BASE;
  shop_generate_inventory(s, r);
  s->shop_reroll_state = (int32_t)(s->reroll_cost * shop_cost_mult(s));
}

/* ------------------------------------------------------------------------------------------
 * round flow  balatro_env_2.py:1326-1381
 * ---------------------------------------------------------------------------------------- */
static void boss_deactivate(BgymState* s) {
  s->boss_type = 0; s->boss_flags = 0; s->boss_cards_required = 0; s->boss_played_types = 0;
  s->boss_hands_played = 0; s->boss_played_cards = 0;
}

static void advance_round(BgymState* s, Rng* r) {
  /* end_of_round_effects returns [] (complete_joker_effects.py:253-259) */
  int gold = 0;
  for (int i = 0; i < s->hand_n; i++) {
    int idx = s->hand[i];
    if (idx < 52 && c16_enh(s->deck[idx]) == BGYM_ENH_GOLD) gold += 3;
  }
  s->money += gold;
  if (s->boss_type) {
    s->money += 5; /* money_reward, boss_blinds.py:60 */
    boss_deactivate(s);
    s->face_down_mask = 0;
  }
  s->round_chips = 0; s->best_hand = 0; s->hands_played_ante = 0;
  if (s->round == 3) {
    s->ante += 1; s->round = 1;
    if (s->ante > 100) return;
  } else {
    s->round += 1;
  }
  s->money += 25 * s->round + (s->round == 3 ? 10 : 0);
  s->hands_left = 4; s->discards_left = 3;
  s->phase = BGYM_PHASE_SHOP;
  generate_shop(s, r);
}

/* ------------------------------------------------------------------------------------------
 * boss blinds  boss_blinds.py:301-507
 * ---------------------------------------------------------------------------------------- */
enum { B_HOOK = 1, B_WALL, B_WHEEL, B_HOUSE, B_MARK, B_FISH, B_PSYCHIC, B_GOAD, B_WATER, B_WINDOW,
       B_MANACLE, B_EYE, B_MOUTH, B_PLANT, B_SERPENT, B_PILLAR, B_NEEDLE, B_HEAD, B_CLUB, B_TOOTH,
       B_FLINT, B_OXIDE, B_ARM, B_VIOLET, B_VERDANT, B_AMBER, B_CRIMSON, B_CERULEAN };

/* can_play_hand :380-407 */
static int boss_can_play(const BgymState* s, int n_played, int ht) {
  switch (s->boss_type) {
    case B_PSYCHIC: return n_played == 5;
    case B_EYE: return !((s->boss_played_types >> ht) & 1);
    case B_MOUTH: return !(s->boss_played_types && !((s->boss_played_types >> ht) & 1));
    case B_VERDANT: return n_played >= s->boss_cards_required;
  }
  return 1;
}

/* modify_scoring :409-445 with _is_card_debuffed :447-478 (suit debuffs compare an IntEnum with a
 * str and never fire — SURVEY Q15) */
static void boss_modify_scoring(const BgymState* s, const int* played_deck_idx, int n_played,
                                int* chips, int* mult) {
  int c = *chips, m = *mult;
  if (s->boss_type == B_FLINT) { c = c / 2; m = m / 2; }
  else if (s->boss_type == B_OXIDE) { c = 0; }
  else if (s->boss_type == B_ARM) { c = (int)(c * 0.75); m = (int)(m * 0.75); }
  int debuffed = 0;
  for (int i = 0; i < n_played; i++) {
    int idx = played_deck_idx[i];
    int rank = code_rank(c16_code(s->deck[idx]));
    int d = 0;
    if (s->boss_type == B_PLANT && rank >= 11 && rank <= 13) d = 1;
    if (s->boss_type == B_VIOLET) d = 1;
    if (s->boss_type == B_PILLAR && ((s->boss_played_cards >> idx) & 1)) d = 1;
    debuffed += d;
  }
  if (debuffed > 0) {
    double penalty = POW_0_8[debuffed];
    c = (int)(c * penalty); m = (int)(m * penalty);
  }
  *chips = c; *mult = m;
}

/* ------------------------------------------------------------------------------------------
 * consumables  balatro_env_2.py:1066-1172, consumables.py:115-655
 * ---------------------------------------------------------------------------------------- */
static void cons_append(BgymState* s, int cid) { if (s->cons_n < 8) s->cons_id[s->cons_n++] = (uint8_t)cid; }
static void cons_pop(BgymState* s, int idx) {
  for (int i = idx; i + 1 < s->cons_n; i++) s->cons_id[i] = s->cons_id[i + 1];
  s->cons_n--;
  s->cons_id[s->cons_n] = 0;
}
static void set_enh(BgymState* s, int idx, int enh) { s->deck[idx] = (uint16_t)((s->deck[idx] & ~(15u << 6)) | ((unsigned)enh << 6)); }
static void set_edition(BgymState* s, int idx, int ed) { s->deck[idx] = (uint16_t)((s->deck[idx] & ~(7u << 10)) | ((unsigned)ed << 10)); }
static void set_seal(BgymState* s, int idx, int seal) { s->deck[idx] = (uint16_t)((s->deck[idx] & ~(7u << 13)) | ((unsigned)seal << 13)); }

static const int PLANET_HAND[12] = {1, 2, 3, 4, 5, 6, 7, 8, 0, 9, 10, 11}; /* balatro_env_2.py:1103-1116 */
static const int WRAITH_JOKER[14] = {BGYM_J_INVISIBLE_JOKER, BGYM_J_BRAINSTORM, BGYM_J_SATELLITE,
  BGYM_J_SHOOT_THE_MOON, 0 /* 'Drivers License' is not a library name */, BGYM_J_CARTOMANCER,
  BGYM_J_ASTRONOMER, BGYM_J_BURNT_JOKER, BGYM_J_BOOTSTRAPS, BGYM_J_CANIO, BGYM_J_TRIBOULET,
  BGYM_J_YORICK, BGYM_J_CHICOT, BGYM_J_PERKEO}; /* consumables.py:479-481 */


/* The deck as the Python list it is in the reference (state.deck): elements in list order.  An element is either one
 * of the original cards.Card objects (equal only to itself: rank + suit equality over 52 distinct cards,
 * cards.py:112-115) or a consumables.Card appended by Cryptid (dataclass equality over rank, suit and the modifiers it
 * was created with, consumables.py:63-70).  card_states stay keyed by list INDEX (balatro_env_2.py:1122-1138). */
typedef struct DeckElem { int code; int appended; int ident; } DeckElem;

static int deck_to_list(const BgymState* s, DeckElem* out) {
  int n = s->deck_n, n_orig = n - s->deck_extra_n;
  for (int i = 0; i < n; i++) {
    if (i < n_orig) { out[i].code = c16_code(s->deck[i]); out[i].appended = 0; out[i].ident = 0; }
    else { out[i].ident = s->deck_extra[i - n_orig]; out[i].code = out[i].ident & 63; out[i].appended = 1; }
  }
  return n;
}

static void list_to_deck(BgymState* s, const DeckElem* list, int n) {
  int n_extra = 0;
  for (int i = 0; i < 52; i++) {                       /* modifiers belong to the index, codes to the list */
    uint16_t mods = (uint16_t)(s->deck[i] & ~63u);
    s->deck[i] = (uint16_t)(mods | (i < n ? list[i].code : 0));
  }
  memset(s->deck_extra, 0, sizeof s->deck_extra);
  for (int i = 0; i < n; i++) if (list[i].appended) s->deck_extra[n_extra++] = (uint16_t)list[i].ident;
  s->deck_extra_n = (uint8_t)n_extra;
  s->deck_n = (uint8_t)n;
}

static int elems_equal(const DeckElem* a, const DeckElem* b) {
  if (a->appended != b->appended) return 0;            /* different classes never compare equal */
  return a->appended ? a->ident == b->ident : a->code == b->code;
}

/* returns reward; *err gets the error code */
static double use_consumable(BgymState* s, int cidx, Rng* r, int* err, int* terminated) {
  *err = BGYM_ERR_NONE;
  int cid = s->cons_id[cidx];
  /* target cards = selected cards in selection order (:1074-1083) */
  int tgt[8], nT = 0;
  for (int k = 0; k < s->sel_n; k++) {
    int sl = sel_slot(s, k);
    if (sl < s->hand_n && s->hand[sl] < s->deck_n) tgt[nT++] = s->hand[sl];
  }
  int success = 0, money_gained = 0, planet_ht = -1, n_affected = 0, n_jokers_created = 0;
  int items[4], n_items = 0, hand_size_change = 0, exception = 0, unsupported = 0;
  int add_jokers[2] = {0, 0}, n_add_jokers = 0, n_created = 0, n_destroyed = 0;
  int tarot = 0;
  if (cid >= 1 && cid <= 22) tarot = cid;
  else if (cid >= 101 && cid <= 122) tarot = cid - 100;

  if (tarot) {
    switch (tarot) {
      case 1: { /* The Fool :127-134 — appends to the aliased state list, no slot check */
        if (s->cons_n > 0) {
          int copied = s->cons_id[rng_below(r, s->cons_n)];
          cons_append(s, copied);
          items[n_items++] = copied; success = 1;
        }
        break;
      }
      case 2: case 4: case 6: { /* Magician / Empress / Hierophant :136-184 */
        int enh = tarot == 2 ? BGYM_ENH_LUCKY : (tarot == 4 ? BGYM_ENH_MULT : BGYM_ENH_BONUS);
        if (nT > 0) {
          for (int i = 0; i < nT && i < 2; i++) { set_enh(s, tgt[i], enh); n_affected++; }
          success = 1;
        }
        break;
      }
      case 7: case 8: case 12: case 16: case 17: { /* Lovers / Chariot / Justice / Devil / Tower */
        int enh = tarot == 7 ? BGYM_ENH_WILD : tarot == 8 ? BGYM_ENH_STEEL : tarot == 12 ? BGYM_ENH_GLASS
                : tarot == 16 ? BGYM_ENH_GOLD : BGYM_ENH_STONE;
        if (nT >= 1) { set_enh(s, tgt[0], enh); n_affected = 1; success = 1; }
        break;
      }
      case 9: /* Strength :202-210 — the rank change is lost */
        if (nT > 0) {
          for (int i = 0; i < nT && i < 2; i++)
            if (code_rank(c16_code(s->deck[tgt[i]])) < 14) n_affected++;
          success = 1;
        }
        break;
      case 10: { /* The Hermit :212-219 */
        int gain = s->money < 20 ? s->money : 20;
        money_gained = gain; success = 1;
        break;
      }
      case 11: /* Wheel of Fortune :221-231 */
        if (nT > 0 && rng_u01(r) < 0.25) {
          int ed = BGYM_ED_FOIL + rng_below(r, 3);
          set_edition(s, tgt[0], ed); n_affected = 1; success = 1;
        }
        break;
      case 13: /* The Hanged Man :241-251 — list.remove of a foreign object raises ValueError */
        if (nT > 0) exception = 1;
        break;
      case 14: /* Death :253-261 — rank/suit copy is lost */
        if (nT >= 2) { n_affected = 2; success = 1; }
        break;
      case 15: { /* Temperance :263-273 */
        int total = 5 * s->joker_n;
        money_gained = total < 50 ? total : 50; success = 1;
        break;
      }
      case 18: case 19: case 20: case 22: /* Star / Moon / Sun / World — suit change is lost */
        if (nT > 0) { n_affected = nT < 3 ? nT : 3; success = 1; }
        break;
      case 3: /* The High Priestess :145-155 */
        for (int i = 0; i < 2; i++) {
          int p = rng_below(r, 9);
          if (s->cons_n < s->cons_slots) { cons_append(s, BGYM_CONS_PLANET_BASE + p); items[n_items++] = BGYM_CONS_PLANET_BASE + p; }
        }
        success = 1;
        break;
      case 5: /* The Emperor :166-175 — creates enum-style names */
        for (int i = 0; i < 2; i++) {
          if (s->cons_n < s->cons_slots) {
            int t = rng_below(r, 22);
            cons_append(s, BGYM_CONS_ENUMSTYLE_BASE + 1 + t); items[n_items++] = BGYM_CONS_ENUMSTYLE_BASE + 1 + t;
          }
        }
        success = 1;
        break;
      case 21: { /* Judgement :318-327 */
        int p = rng_below(r, 9);
        if (s->cons_n < s->cons_slots) { cons_append(s, BGYM_CONS_PLANET_BASE + p); items[n_items++] = BGYM_CONS_PLANET_BASE + p; }
        success = 1;
        break;
      }
    }
  } else if (cid >= 30 && cid <= 41) { /* planets :644-653 */
    success = 1; planet_ht = PLANET_HAND[cid - 30];
  } else if (cid >= 50 && cid <= 67) {
    switch (cid - 50) {
      case 0: case 1: case 2: /* Familiar / Grim / Incantation: deck.remove(ad-hoc class) raises */
        if (nT >= 1) exception = 1;
        break;
      case 3: case 11: case 13: case 14: { /* Talisman / Deja Vu / Trance / Medium: consumables.Seal
        values (RED1 BLUE2 GOLD3 PURPLE4, consumables.py:56-61) stored raw and later read as
        cards.Seal (GOLD1 RED2 BLUE3 PURPLE4) — SURVEY Q18 */
        int v = (cid - 50) == 3 ? 3 : (cid - 50) == 11 ? 1 : (cid - 50) == 13 ? 2 : 4;
        if (nT >= 1) { set_seal(s, tgt[0], v); n_affected = 1; success = 1; }
        break;
      }
      case 4: /* Aura :468-475 */
        if (nT >= 1) { set_edition(s, tgt[0], BGYM_ED_FOIL + rng_below(r, 3)); n_affected = 1; success = 1; }
        break;
      case 5: /* Wraith :477-489 */
        if (s->joker_n < s->joker_slots) {
          int j = WRAITH_JOKER[rng_below(r, 14)];
          add_jokers[n_add_jokers++] = j; n_jokers_created = 1; hand_size_change = -1; success = 1;
        }
        break;
      case 6: case 7: /* Sigil / Ouija: assigning to a frozen Card raises FrozenInstanceError */
        if (s->hand_n > 0) { (void)rng_below(r, (cid - 50) == 6 ? 4 : 13); exception = 1; }
        break;
      case 8: /* Ectoplasm :512-518 */
        if (s->joker_n > 0) { hand_size_change = -1; success = 1; }
        break;
      case 10: /* Ankh :534-544 — the created "joker" is a dict, never added; still rewarded */
        if (s->joker_n > 0) { (void)rng_below(r, s->joker_n); n_jokers_created = 1; success = 1; }
        break;
      case 12: /* Hex :554-564 */
        if (s->joker_n > 0) { (void)rng_below(r, s->joker_n); success = 1; }
        break;
      case 16: /* The Soul :594-602 */
        if (s->joker_n < s->joker_slots) {
          add_jokers[n_add_jokers++] = BGYM_J_CANIO + rng_below(r, 5); n_jokers_created = 1; success = 1;
        }
        break;
      case 17: success = 1; break; /* Black Hole :604-611 */
      case 9: { /* Immolate :520-532 — random.sample(deck, min(5, len(deck))), then deck.remove(card) one by one */
        DeckElem list[64], victims[5];
        int n = deck_to_list(s, list);
        int k = n < 5 ? n : 5, picked[5];
        rng_sample(r, n, k, picked);
        for (int t = 0; t < k; t++) victims[t] = list[picked[t]];
        /* the blind's played_cards set holds id(card): it follows the card objects through the removals */
        int pillar[64];
        for (int i = 0; i < n; i++) pillar[i] = (int)((s->boss_played_cards >> i) & 1);
        for (int t = 0; t < k; t++) {
          int at = 0;
          while (!elems_equal(&list[at], &victims[t])) at++;   /* list.remove: first equal element */
          for (int i = at; i + 1 < n; i++) { list[i] = list[i + 1]; pillar[i] = pillar[i + 1]; }
          n--;
        }
        list_to_deck(s, list, n);
        s->boss_played_cards = 0;
        for (int i = 0; i < n; i++) if (pillar[i]) s->boss_played_cards |= 1ull << i;
        money_gained = 20; n_destroyed = k; success = 1;
        break;
      }
      case 15: /* Cryptid :582-592 — two consumables.Card copies of the first target go to the end of the deck list */
        if (nT >= 1) {
          if (s->deck_extra_n + 2 > 4) { unsupported = 1; break; }  /* capacity of BgymHot.deck_extra (include/bgym.h) */
          DeckElem list[64];
          int n = deck_to_list(s, list);
          for (int q = 0; q < 2; q++) {
            list[n].code = c16_code(s->deck[tgt[0]]); list[n].appended = 1; list[n].ident = s->deck[tgt[0]];
            n++;
          }
          list_to_deck(s, list, n);
          n_created = 2; success = 1;
        }
        break;
      default: break;
    }
  }

  if (exception) { /* SafeBalatroEnv convention train_balatro_fixed.py:262-269 */
    *err = BGYM_ERR_REF_EXCEPTION; *terminated = 1;
    return -100.0;
  }
  double reward = 0.0;
  if (unsupported) { *err = BGYM_ERR_UNSUPPORTED; reward = -1.0; }
  else if (success) {
    cons_pop(s, cidx);                                                   /* :1094 */
    if (money_gained > 0) { s->money += money_gained; reward += money_gained / 10.0; }
    if (planet_ht >= 0) {                                                /* :1101-1120 */
      if (s->hand_level[planet_ht] < 255) s->hand_level[planet_ht]++;
      reward += 10.0;
    }
    if (n_affected > 0) reward += n_affected * 2.0;                      /* :1122-1138 */
    if (n_created > 0) reward += n_created * 3.0;                        /* :1140-1141 */
    if (n_destroyed > 0) reward += n_destroyed * 1.0;                    /* :1143-1144 */
    if (n_jokers_created > 0) {                                          /* :1146-1154 */
      for (int i = 0; i < n_add_jokers; i++)
        if (s->joker_n < s->joker_slots && add_jokers[i] != 0 && s->joker_n < 8) s->joker_id[s->joker_n++] = (uint8_t)add_jokers[i];
      reward += n_jokers_created * 15.0;
    }
    if (n_items > 0) {                                                   /* :1156-1160 */
      for (int i = 0; i < n_items; i++)
        if (s->cons_n < s->cons_slots) cons_append(s, items[i]);
      reward += n_items * 5.0;
    }
    if (hand_size_change) {                                             /* :1162-1164 */
      int hs = (int)s->hand_size + hand_size_change;
      s->hand_size = (uint8_t)(hs < 0 ? 0 : hs); /* <= 0 draws nothing either way */
    }
  } else {
    reward = -1.0; *err = BGYM_ERR_CONSUMABLE_FAILED;                    /* :1167-1169 */
  }
  s->sel_n = 0; s->sel_order = 0;                                        /* :1171 */
  return reward;
}

/* ------------------------------------------------------------------------------------------
 * step  balatro_env_2.py:616-1064, 1174-1318
 * ---------------------------------------------------------------------------------------- */
static int joker_named(const BgymState* s, int id) { return joker_owned(s, id); }

static void step_env(BgymState* s, int action, const BgymDraws* tape, double* reward_out,
                     uint8_t* term_out, BgymInfo* info) {
  Rng rng;
  rng_init(&rng, s, tape);
  double reward = 0.0;
  int terminated = 0;
  memset(info, 0, sizeof *info);
  info->hand_type = -1; info->x_mult = 1.0;

  /* terminal guards :619-623 */
  if (s->ante > 100 || s->chips_scored > 1000000000LL) {
    info->flags |= BGYM_F_GUARD_TERMINATED;
    *reward_out = 0.0; *term_out = 1;
    return;
  }
  /* mask validation :626-627 */
  if (action < 0 || action >= BGYM_NUM_ACTIONS || !((action_mask(s) >> action) & 1)) {
    info->error_code = BGYM_ERR_INVALID_ACTION;
    *reward_out = -1.0; *term_out = 0;
    return;
  }
  s->ep_len++;

  if (s->phase == BGYM_PHASE_PLAY) {
    if (action == BGYM_A_PLAY_HAND) {
      /* selected cards -> real cards :650-660 */
      int played_idx[8], n_played = 0;
      ScoreCard sc[8];
      for (int k = 0; k < s->sel_n; k++) {
        int sl = sel_slot(s, k);
        if (sl < s->hand_n && s->hand[sl] < s->deck_n) {
          int idx = s->hand[sl];
          uint16_t c = s->deck[idx];
          played_idx[n_played] = idx;
          sc[n_played].chip_value = card_chip_value(c16_code(c), c16_enh(c), c16_edition(c));
          sc[n_played].rank = code_rank(c16_code(c)); sc[n_played].suit = code_suit(c16_code(c));
          n_played++;
        }
      }
      /* highlight :663-666 — slots accumulate until the next discard */
      for (int k = 0; k < s->sel_n; k++) {
        int sl = sel_slot(s, k);
        if (sl < s->hand_n) s->highlight_mask |= (uint8_t)(1u << sl);
      }
      /* classify deck[slot] for every highlighted slot :669-671 */
      int codes[8], nc = 0;
      for (int sl = 0; sl < 8; sl++)
        if ((s->highlight_mask >> sl) & 1) codes[nc++] = c16_code(s->deck[sl]);
      int ht = classify_hand(codes, nc);
      /* boss gate :677-680 */
      if (s->boss_type && !boss_can_play(s, n_played, ht)) {
        info->error_code = BGYM_ERR_BOSS_RESTRICTION;
        *reward_out = -1.0; *term_out = 0;
        return;
      }
      /* UnifiedScorer.score_hand with game_state['jokers'] = list of dicts -> no joker fires */
      ScoreIn in;
      memset(&in, 0, sizeof in);
      in.cards = sc; in.n_cards = n_played; in.jokers = NULL; in.n_jokers = 0; in.hand_type = ht;
      ScoreOut so;
      score_hand(&in, &so, s->hand_level);
      int64_t base_score = so.score;
      /* per-card enhancement / seal loop :703-734 */
      int extra_money = 0, n_red = 0, planets[8], n_planets = 0;
      for (int i = 0; i < n_played; i++) {
        uint16_t c = s->deck[played_idx[i]];
        int enh = c16_enh(c), seal = c16_seal(c);
        if (enh == BGYM_ENH_GLASS) {
          (void)rng_u01(&rng);                       /* break roll, result unused (:712-713, :770-772) */
        } else if (enh == BGYM_ENH_LUCKY) {
          double mult_roll = rng_u01(&rng);          /* +20 mult computed then dropped (:721-722, :738) */
          double money_roll = rng_u01(&rng);
          (void)mult_roll;
          if (money_roll < 0.0667) extra_money += 20;
        }
        if (seal == BGYM_SEAL_GOLD) extra_money += 3;
        else if (seal == BGYM_SEAL_RED) n_red++;
        else if (seal == BGYM_SEAL_BLUE) {
          /* planet for this hand type (cards.py:228-246); room is tested against the CURRENT list */
          static const int HT_PLANET[12] = {38, 30, 31, 32, 33, 34, 35, 36, 37, 39, 40, 41};
          if (s->cons_n < s->cons_slots) planets[n_planets++] = HT_PLANET[ht];
        }
      }
      int64_t final_score = base_score;
      /* steel cards held in hand :560-570, :741-742 */
      double steel = 1.0;
      for (int i = 0; i < s->hand_n; i++) {
        int idx = s->hand[i], selected = 0;
        for (int k = 0; k < n_played; k++) if (played_idx[k] == idx) selected = 1;
        /* the selected set is built from hand slots < len(hand_indexes) without the deck check */
        for (int k = 0; k < s->sel_n; k++) { int sl = sel_slot(s, k); if (sl < s->hand_n && s->hand[sl] == idx) selected = 1; }
        if (!selected && idx < 52 && c16_enh(s->deck[idx]) == BGYM_ENH_STEEL) steel *= 1.5;
      }
      final_score = (int64_t)((double)final_score * steel);
      /* boss modification ratio :745-755 */
      if (s->boss_type) {
        int bc, bm;
        hand_chips_mult(s->hand_level, ht, &bc, &bm);
        int mc = bc, mm = bm;
        boss_modify_scoring(s, played_idx, n_played, &mc, &mm);
        if (bc > 0 && bm > 0) {
          double chip_ratio = (double)mc / (double)bc;
          double mult_ratio = (double)mm / (double)bm;
          final_score = (int64_t)((double)final_score * chip_ratio * mult_ratio);
        }
      }
      /* retriggers :757-759 */
      double retrigger_bonus = n_red * 0.5;
      final_score = (int64_t)((double)final_score * (1 + retrigger_bonus));
      s->money += extra_money;
      for (int i = 0; i < n_planets; i++)
        if (s->cons_n < s->cons_slots) cons_append(s, planets[i]);

      int64_t old_round = s->round_chips;
      double needed = (double)(s->chips_needed > 1 ? s->chips_needed : 1);
      double old_progress = fmin(1.0, (double)old_round / needed);
      s->round_chips += final_score;
      s->chips_scored += final_score;
      s->hands_played_total += 1;
      s->hands_played_ante += 1;
      if (final_score > s->best_hand) s->best_hand = clamp_i32(final_score);
      if (s->hand_play_count[ht] < 255) s->hand_play_count[ht]++;
      /* boss on_hand_scored :480-507 (Tooth / Serpent write into a throw-away dict — SURVEY Q15) */
      if (s->boss_type) {
        s->boss_played_types |= (uint16_t)(1u << ht);
        s->boss_flags &= (uint8_t)~1u;
        s->boss_hands_played++;
        if (s->boss_type == B_PILLAR)
          for (int i = 0; i < n_played; i++) s->boss_played_cards |= 1ull << played_idx[i];
        if (s->boss_type == B_VERDANT && s->boss_cards_required < 7) s->boss_cards_required++;
      }
      s->sel_n = 0; s->sel_order = 0;

      /* reward shaping :799-892 */
      double new_progress = fmin(1.0, (double)s->round_chips / needed);
      double progress_reward = 15.0 * new_progress;
      double milestone = 0.0;
      if (old_progress < 0.25 && 0.25 <= new_progress) milestone = 5.0;
      else if (old_progress < 0.5 && 0.5 <= new_progress) milestone = 10.0;
      else if (old_progress < 0.75 && 0.75 <= new_progress) milestone = 15.0;
      else if (old_progress < 1.0 && 1.0 <= new_progress) milestone = 25.0;
      double score_reward;
      if (s->ante <= 3) score_reward = fmin(10.0, (double)final_score / 100.0);
      else score_reward = fmin(10.0, 3.0 * log10((double)(final_score > 1 ? final_score : 1)));
      static const double HQ[12] = {0.1, 0.5, 1.0, 2.0, 2.5, 2.5, 3.5, 5.0, 7.0, 10.0, 0.0, 0.0};
      double hand_quality = HQ[ht];
      double efficiency = 0.0;
      if (ht >= BGYM_HT_THREE_KIND && n_played <= 3) efficiency = 2.0;
      else if (ht >= BGYM_HT_FLUSH && n_played == 5) efficiency = 1.0;
      else if (n_played <= 4 && s->hands_left <= 2) efficiency = 1.5;
      double synergy = 0.0;
      if (ht == BGYM_HT_FLUSH && (joker_named(s, BGYM_J_SMEARED_JOKER) || joker_named(s, BGYM_J_FOUR_FINGERS) ||
                                  joker_named(s, BGYM_J_SHORTCUT))) synergy += 2.0;
      if ((ht == BGYM_HT_ONE_PAIR || ht == BGYM_HT_TWO_PAIR || ht == BGYM_HT_THREE_KIND) &&
          (joker_named(s, BGYM_J_ODD_TODD) || joker_named(s, BGYM_J_EVEN_STEVEN) ||
           joker_named(s, BGYM_J_JOLLY_JOKER) || joker_named(s, BGYM_J_ZANY_JOKER))) synergy += 1.5;
      int face_cards = 0;
      for (int i = 0; i < n_played; i++) if (sc[i].rank >= 11) face_cards++;   /* J Q K and A (:862) */
      if (face_cards > 0 && (joker_named(s, BGYM_J_SCARY_FACE) || joker_named(s, BGYM_J_SMILEY_FACE) ||
                             joker_named(s, BGYM_J_BUSINESS_CARD))) synergy += 0.5 * face_cards;
      double strategy = 0.0;
      if (new_progress > 0.7 && s->hands_left >= 3) strategy = 2.0;
      else if (new_progress < 0.3 && ht >= BGYM_HT_FLUSH) strategy = 3.0;
      double ante_bonus = 0.0;
      if (s->ante >= 4) ante_bonus = fmin(5.0, (s->ante - 3) * 0.5);
      reward = progress_reward + milestone + score_reward + hand_quality * 2.0 + efficiency * 1.5 +
               synergy * 3.0 + strategy * 2.0 + ante_bonus;
      reward = fmin(reward, 100.0);

      info->final_score = final_score; info->base_score = clamp_i32(base_score);
      info->chips = so.chips; info->mult = so.mult; info->x_mult = so.x_mult;
      info->hand_type = (int8_t)ht; info->cards_played = (uint8_t)n_played; info->flags |= BGYM_F_PLAYED;

      /* round end :914-960 */
      if (s->round_chips >= s->chips_needed) {
        double bonus = 25.0 + (10.0 * s->ante);
        reward += fmin(50.0, bonus);
        advance_round(s, &rng);
        info->flags |= BGYM_F_BEAT_BLIND;
      } else if (s->hands_left <= 1) {
        reward += -50.0 * (1.0 - new_progress);
        terminated = 1;
        info->flags |= BGYM_F_FAILED;
      } else {
        s->hands_left -= 1;
        draw_cards(s);
        if (s->boss_type) { /* on_hand_drawn boss_blinds.py:343-378; runs with first_hand already False */
          int face = 0;
          if (s->boss_type == B_HOOK) {
            if (s->hand_n >= 2) {
              int pick[2];
              rng_sample(&rng, s->hand_n, 2, pick);
              int hi = pick[0] > pick[1] ? pick[0] : pick[1], lo = pick[0] > pick[1] ? pick[1] : pick[0];
              hand_pop(s, hi); hand_pop(s, lo);
            }
          } else if (s->boss_type == B_WHEEL) {
            for (int i = 0; i < s->hand_n; i++) if (rng_u01(&rng) < 1.0 / 7) face |= 1 << i;
          } else if (s->boss_type == B_MARK) {
            for (int i = 0; i < s->hand_n; i++) {
              int rank = code_rank(c16_code(s->deck[s->hand[i]]));
              if (rank >= 11 && rank <= 13) face |= 1 << i;
            }
          } else if (s->boss_type == B_FISH) {
            face = (1 << s->hand_n) - 1;
          }
          s->face_down_mask = (uint8_t)face;
        }
      }
    } else if (action == BGYM_A_DISCARD) {
      /* :962-1050 */
      int n_disc = 0, purple = 0, faces = 0;
      for (int k = 0; k < s->sel_n; k++) {
        int sl = sel_slot(s, k);
        if (sl < s->hand_n && s->hand[sl] < s->deck_n) {
          uint16_t c = s->deck[s->hand[sl]];
          if (c16_seal(c) == BGYM_SEAL_PURPLE) purple++;
          int rank = code_rank(c16_code(c));
          if (rank >= 11 && rank <= 13) faces++;
          n_disc++;
        }
      }
      int is_first = s->discards_left == 3; /* game.discards is the constant 3 (balatro_game.py:25) */
      int money_from_discards = 0, n_discard_jokers = 0;
      for (int j = 0; j < s->joker_n; j++) {
        int id = s->joker_id[j], money = 0;
        if (id == BGYM_J_TRADING_CARD && is_first && n_disc == 1) money = 3;        /* complete_joker_effects.py:189-191 */
        else if (id == BGYM_J_FACELESS_JOKER && faces >= 3) money = 5;            /* :193-197 */
        money_from_discards += money; s->money += money;
        if (id == BGYM_J_FACELESS_JOKER || id == BGYM_J_HIT_THE_ROAD || id == BGYM_J_RESERVED_PARKING ||
            id == BGYM_J_LUCHADOR) n_discard_jokers++;
      }
      /* highlight then discard_hand balatro_game.py:111-127 — stale highlights go too */
      for (int k = 0; k < s->sel_n; k++) { int sl = sel_slot(s, k); if (sl < s->hand_n) s->highlight_mask |= (uint8_t)(1u << sl); }
      for (int sl = 7; sl >= 0; sl--)
        if (((s->highlight_mask >> sl) & 1) && sl < s->hand_n) hand_pop(s, sl);
      s->highlight_mask = 0;
      draw_cards(s);
      s->discards_left -= 1;
      s->sel_n = 0; s->sel_order = 0;
      /* purple seals -> tarots :1021-1032 (rng stream 'seal_applications') */
      for (int i = 0; i < purple; i++)
        if (s->cons_n < s->cons_slots) cons_append(s, BGYM_CONS_TAROT_BASE + rng_below(&rng, 22));
      reward = 0.2;
      if (n_discard_jokers) reward += 0.5 * n_discard_jokers;
      if (money_from_discards > 0) reward += money_from_discards / 5.0;
      double progress = (double)s->round_chips / (double)(s->chips_needed > 1 ? s->chips_needed : 1);
      if (progress < 0.5 && s->discards_left > 1) reward += 0.5;
      else if (progress > 0.8 && s->discards_left > 1) reward -= 0.3;
    } else if (action >= BGYM_A_SELECT_BASE && action < BGYM_A_SELECT_BASE + 8) {
      /* toggle :1052-1058 — ordered list, no 5-card cap */
      int slot = action - BGYM_A_SELECT_BASE;
      if (slot < s->hand_n) {
        int found = -1;
        for (int k = 0; k < s->sel_n; k++) if (sel_slot(s, k) == slot) found = k;
        if (found >= 0) {
          uint32_t lo = s->sel_order & ((1u << (4 * found)) - 1);
          uint32_t hi = found == 7 ? 0 : (s->sel_order >> (4 * (found + 1))) << (4 * found);
          s->sel_order = lo | hi;
          s->sel_n--;
        } else {
          s->sel_order |= (uint32_t)slot << (4 * s->sel_n);
          s->sel_n++;
        }
      }
    } else if (action >= BGYM_A_USE_CONS_BASE && action < BGYM_A_USE_CONS_BASE + 5) {
      int err = 0;
      reward = use_consumable(s, action - BGYM_A_USE_CONS_BASE, &rng, &err, &terminated);
      info->error_code = (uint8_t)err;
    }
  } else if (s->phase == BGYM_PHASE_SHOP) {
    /* :1174-1253 */
    if (action == BGYM_A_SHOP_END) {
      s->phase = BGYM_PHASE_PLAY;
      draw_cards(s);
      info->flags |= BGYM_F_SHOP_DONE;
      reward = 0.0;
    } else if (action == BGYM_A_SHOP_REROLL) {
      int cost = (int)(s->reroll_cost * shop_cost_mult(s));          /* shop.py:172 */
      if (s->money < cost) { reward = -1.0; info->error_code = BGYM_ERR_SHOP; }
      else {
        s->money -= cost;
        s->reroll_cost = (int32_t)(s->reroll_cost * 1.35);
        shop_generate_inventory(s, &rng);
        reward = 0.0;
      }
    } else if (action >= BGYM_A_SHOP_BUY_BASE && action < BGYM_A_SHOP_BUY_BASE + 10) {
      int i = action - BGYM_A_SHOP_BUY_BASE;
      int type = s->item_type[i], id = s->item_id[i], cost = s->item_cost[i];
      s->money -= cost;                                              /* shop.py:185-187 */
      for (int k = i; k + 1 < s->n_items; k++) {
        s->item_type[k] = s->item_type[k + 1]; s->item_id[k] = s->item_id[k + 1]; s->item_cost[k] = s->item_cost[k + 1];
      }
      s->n_items--;
      s->item_type[s->n_items] = 0; s->item_id[s->n_items] = 0; s->item_cost[s->n_items] = 0;
      if (type == BGYM_ITEM_PACK) {
        int count = id == BGYM_PACK_STANDARD ? 3 : 1;                /* shop.py:150-157 */
        for (int k = 0; k < count; k++) (void)rng_below(&rng, 52);
        reward = 5.0;
      } else if (type == BGYM_ITEM_CARD) {
        reward = 3.0;
      } else if (type == BGYM_ITEM_JOKER) {
        if (s->joker_n >= 5) { reward = -1.0; info->error_code = BGYM_ERR_SHOP; } /* shop.py:196-197 */
        else { s->joker_id[s->joker_n++] = (uint8_t)id; reward = 15.0; }
      } else if (type == BGYM_ITEM_VOUCHER) {
        if (id == BGYM_VOUCHER_MAGIC_TRICK) s->n_magic_trick++; else s->n_minimalist++;
        reward = 10.0;
      }
    } else if (action >= BGYM_A_SELL_JOKER_BASE && action < BGYM_A_SELL_JOKER_BASE + 5) {
      int j = action - BGYM_A_SELL_JOKER_BASE;
      int id = s->joker_id[j];
      for (int k = j; k + 1 < s->joker_n; k++) s->joker_id[k] = s->joker_id[k + 1];
      s->joker_n--;
      s->joker_id[s->joker_n] = 0;
      int sell = JOKER_COST[id] / 2;
      if (sell < 3) sell = 3;
      s->money += sell; s->jokers_sold++;
      reward = sell / 5.0;
    }
  } else if (s->phase == BGYM_PHASE_BLIND_SELECT) {
    /* :1255-1318 */
    if (action >= BGYM_A_SELECT_BLIND_BASE && action < BGYM_A_SELECT_BLIND_BASE + 3) {
      int bt = action - BGYM_A_SELECT_BLIND_BASE;
      s->round = (uint8_t)(bt + 1);
      int64_t needed;
      if (s->ante <= 8) needed = BLIND_CHIPS[s->ante - 1][bt];
      else needed = (int64_t)(BLIND_CHIPS[7][bt] * POW_1_5[s->ante - 8]);
      if (bt == 2) {
        int boss = 1 + rng_below(&rng, 28);            /* select_boss_blind boss_blinds.py:522-532 */
        s->boss_type = (uint8_t)boss;                  /* activate_boss_blind :308-341 */
        s->boss_flags = 1; s->boss_cards_required = 5; s->boss_played_types = 0;
        s->boss_hands_played = 0; s->boss_played_cards = 0;
        double chip_mult = boss == B_WALL ? 2.0 : 1.0;
        needed = (int64_t)((double)needed * chip_mult);
        if (boss == B_WATER) s->discards_left = 0;
        if (boss == B_MANACLE) s->hand_size -= 1;
        if (boss == B_NEEDLE) s->hands_left = 1;
        reward = 10.0;
      }
      s->chips_needed = clamp_i32(needed);
      s->phase = BGYM_PHASE_PLAY;
      draw_cards(s);
    } else if (action == BGYM_A_SKIP_BLIND) {
      reward = -5.0;
      advance_round(s, &rng);
    }
  }
  refresh_hand_codes(s);
  *reward_out = reward;
  *term_out = (uint8_t)terminated;
}

int oracle_step(BgymState* state, int32_t* actions, const BgymDraws* draws, BgymObs* obs,
                double* reward, uint8_t* terminated, uint8_t* truncated, BgymInfo* info,
                int64_t n, int flags) {
  for (int64_t i = 0; i < n; i++) {
    BgymInfo inf;
    if (flags & BGYM_FLAG_RANDOM_POLICY) {
      /* uniform legal action: Philox keyed (rng_seed, policy key), counter = steps in the episode */
      uint64_t m = action_mask(&state[i]);
      int cnt = __builtin_popcountll(m), act = 0;
      if (cnt) {
        uint32_t w[4];
        philox4x32_10(state[i].ep_len, 0, 0, 0, state[i].rng_seed, 0x5A17AC71u, w);
        int k = (int)(((uint64_t)w[0] * (uint64_t)cnt) >> 32);
        for (int t = 0; t < k; t++) m &= m - 1;
        act = __builtin_ctzll(m);
      }
      actions[i] = act;
    }
    step_env(&state[i], actions[i], draws ? &draws[i] : NULL, &reward[i], &terminated[i], &inf);
    if (truncated) truncated[i] = 0;
    if (terminated[i] && (flags & BGYM_FLAG_AUTORESET)) {
      uint32_t episode = state[i].episode + 1;
      reset_env(&state[i], next_episode_seed(state[i].rng_seed), NULL, flags);
      state[i].episode = episode;
      inf.flags |= BGYM_F_AUTORESET_DONE;
    }
    if (info) info[i] = inf;
    if (obs && !(flags & BGYM_FLAG_NO_OBS)) write_obs(&state[i], &obs[i]);
  }
  return 0;
}

/* uniform random legal action from the mask word: Philox keyed by (seed, env index, step) */
int oracle_sample_actions(const BgymObs* obs, int32_t* actions, uint32_t seed, uint64_t step, int64_t n) {
  for (int64_t i = 0; i < n; i++) {
    uint64_t m = obs[i].action_mask_bits;
    int cnt = __builtin_popcountll(m);
    if (cnt == 0) { actions[i] = 0; continue; }
    uint32_t w[4];
    philox4x32_10((uint32_t)i, (uint32_t)((uint64_t)i >> 32), (uint32_t)step, (uint32_t)(step >> 32), seed,
                  0x5A17AC71u, w);
    int k = (int)(((uint64_t)w[0] * (uint64_t)cnt) >> 32);
    for (int a = 0; a < 64; a++) {
      if (!((m >> a) & 1)) continue;
      if (k == 0) { actions[i] = a; break; }
      k--;
    }
  }
  return 0;
}

int oracle_sizes(int* out) {
  out[0] = (int)sizeof(BgymState); out[1] = (int)sizeof(BgymObs); out[2] = (int)sizeof(BgymInfo);
  out[3] = (int)sizeof(BgymDraws); out[4] = (int)sizeof(BgymScoreCtx);
  return 0;
}
