// CTC prefix beam search on the host (the reference's other Decoder: decoder.py:147-231, `prefix_beam_search`, behind
// PrefixBeamSearchLMDecoder.decode, decoder.py:233-267).  The reference keeps two dict-of-Counter tables keyed by Python strings
// and re-sorts all candidate prefixes every frame; here prefixes are nodes of a trie, the two probability tables are flat arrays
// with per-frame stamps, and a batch is spread over host threads.  The arithmetic (float64, same association order), the
// candidate ORDER (Counter insertion order, stable descending sort) and the pruning rules are reproduced exactly, so transcripts
// are bit-identical to the reference's whenever no language-model callback is involved.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/w2l_sm100.h"

namespace {

struct Node {
  int32_t parent;
  int32_t ch;            // label id appended to the parent (-1 for the root)
  int32_t len;
  int32_t words;         // matches of r'\w+[\s|>]' in the prefix
  uint8_t last_is_word;  // last character matches \w
  uint8_t has_nonspace;  // len(l.replace(' ', '')) > 0
  uint8_t ends;          // last character is end_char
};

struct Search {
  const double* probs;
  int64_t T, F;
  int32_t blank, space_id, end_id, k;
  double alpha, beta, prune;
  const uint8_t* is_word;   // per label: matches \w
  const uint8_t* is_term;   // per label: matches [\s|>]
  w2l_lm_callback lm;
  void* lm_user;

  std::vector<Node> nodes;
  std::vector<std::vector<std::pair<int32_t, int32_t>>> children;   // (label, child) per node: alphabets are tiny
  // Pb/Pnb of the previous and the current frame; stamp == frame means "key present in that Counter"
  std::vector<double> pb_prev, pnb_prev, pb_cur, pnb_cur;
  std::vector<int64_t> sb_prev, snb_prev, sb_cur, snb_cur;
  std::vector<int32_t> order_b, order_nb;    // insertion order of the current frame's Counters
  std::vector<int64_t> in_beam;              // stamp: node is in A_prev of the current frame

  int32_t child(int32_t n, int32_t c) {
    for (auto& e : children[n])
      if (e.first == c) return e.second;
    Node nd;
    const Node& p = nodes[n];
    nd.parent = n;
    nd.ch = c;
    nd.len = p.len + 1;
    nd.words = p.words + ((is_term[c] && p.last_is_word) ? 1 : 0);
    nd.last_is_word = is_word[c];
    nd.has_nonspace = p.has_nonspace || c != space_id;
    nd.ends = c == end_id;
    const int32_t id = (int32_t)nodes.size();
    nodes.push_back(nd);
    children.emplace_back();
    children[n].emplace_back(c, id);
    for (auto* v : {&pb_prev, &pnb_prev, &pb_cur, &pnb_cur}) v->push_back(0.0);
    for (auto* v : {&sb_prev, &snb_prev, &sb_cur, &snb_cur, &in_beam}) v->push_back(-1);
    return id;
  }
  double get_prev_b(int32_t n, int64_t t) const { return sb_prev[n] == t - 1 ? pb_prev[n] : 0.0; }
  double get_prev_nb(int32_t n, int64_t t) const { return snb_prev[n] == t - 1 ? pnb_prev[n] : 0.0; }
  void add_b(int32_t n, int64_t t, double v) {          // Pb[t][n] += v   (Counter: a missing key counts as int 0)
    if (sb_cur[n] != t) {
      sb_cur[n] = t;
      pb_cur[n] = 0.0;
      order_b.push_back(n);
    }
    pb_cur[n] = pb_cur[n] + v;
  }
  void add_nb(int32_t n, int64_t t, double v) {
    if (snb_cur[n] != t) {
      snb_cur[n] = t;
      pnb_cur[n] = 0.0;
      order_nb.push_back(n);
    }
    pnb_cur[n] = pnb_cur[n] + v;
  }
  void set_b(int32_t n, int64_t t, double v) {
    if (sb_cur[n] != t) {
      sb_cur[n] = t;
      order_b.push_back(n);
    }
    pb_cur[n] = v;
  }
  void set_nb(int32_t n, int64_t t, double v) {
    if (snb_cur[n] != t) {
      snb_cur[n] = t;
      order_nb.push_back(n);
    }
    pnb_cur[n] = v;
  }
  void ids_of(int32_t n, std::vector<int32_t>& out) const {
    out.resize(nodes[n].len);
    for (int32_t i = nodes[n].len - 1; i >= 0; --i) {
      out[i] = nodes[n].ch;
      n = nodes[n].parent;
    }
  }
  double word_bonus(int32_t n) const {                   // (len(W(l)) + 1) ** beta
    const double base = (double)(nodes[n].words + 1);
    if (beta == std::floor(beta) && beta >= 0 && beta <= 64) {   // int ** int in Python: exact product
      double r = 1.0;
      for (int i = 0; i < (int)beta; ++i) r *= base;
      return r;
    }
    return std::pow(base, beta);
  }
  double lm_factor(int32_t l_plus) {
    if (!lm) return 1.0;                                   // (lambda l: 1)(...) ** alpha == 1
    std::vector<int32_t> ids;
    ids_of(l_plus, ids);
    size_t a = 0, b = ids.size();                          // l_plus.strip(' ' + end_char)
    while (a < b && (ids[a] == space_id || ids[a] == end_id)) ++a;
    while (b > a && (ids[b - 1] == space_id || ids[b - 1] == end_id)) --b;
    return std::pow(lm(ids.data() + a, (int64_t)(b - a), lm_user), alpha);
  }

  // returns the best prefix's node and its score
  int32_t run(double* score_out) {
    nodes.clear();
    children.clear();
    Node root;
    root.parent = -1;
    root.ch = -1;
    root.len = 0;
    root.words = 0;
    root.last_is_word = 0;
    root.has_nonspace = 0;
    root.ends = 0;
    nodes.push_back(root);
    children.emplace_back();
    for (auto* v : {&pb_prev, &pnb_prev, &pb_cur, &pnb_cur}) v->assign(1, 0.0);
    for (auto* v : {&sb_prev, &snb_prev, &sb_cur, &snb_cur, &in_beam}) v->assign(1, -1);
    pb_prev[0] = 1.0;                                      // Pb[0][''] = 1, Pnb[0][''] = 0
    pnb_prev[0] = 0.0;
    sb_prev[0] = snb_prev[0] = 0;
    std::vector<int32_t> beam{0}, alphabet, cand;
    std::vector<double> cand_val, cand_key;
    std::vector<int32_t> perm;
    for (int64_t t = 1; t <= T; ++t) {                     // the reference prepends an all-zero frame: its t runs 1..T
      const double* row = probs + (t - 1) * F;
      alphabet.clear();
      for (int32_t c = 0; c < F; ++c)
        if (row[c] > prune) alphabet.push_back(c);
      order_b.clear();
      order_nb.clear();
      for (int32_t l : beam) in_beam[l] = t;
      for (size_t bi = 0; bi < beam.size(); ++bi) {
        const int32_t l = beam[bi];
        const double pb_l = get_prev_b(l, t), pnb_l = get_prev_nb(l, t);
        if (nodes[l].len > 0 && nodes[l].ends) {
          set_b(l, t, pb_l);
          set_nb(l, t, pnb_l);
          continue;
        }
        for (int32_t c : alphabet) {
          if (c == blank) {
            add_b(l, t, row[blank] * (pb_l + pnb_l));
            continue;
          }
          const int32_t lp = child(l, c);
          if (nodes[l].len > 0 && c == nodes[l].ch) {
            add_nb(lp, t, row[c] * pb_l);
            add_nb(l, t, row[c] * pnb_l);
          } else if (nodes[l].has_nonspace && (c == space_id || c == end_id)) {
            const double lm_prob = lm_factor(lp);
            add_nb(lp, t, lm_prob * row[c] * (pb_l + pnb_l));
          } else {
            add_nb(lp, t, row[c] * (pb_l + pnb_l));
          }
          if (in_beam[lp] != t) {                          // make use of discarded prefixes
            add_b(lp, t, row[blank] * (get_prev_b(lp, t) + get_prev_nb(lp, t)));
            add_nb(lp, t, row[c] * get_prev_nb(lp, t));
          }
        }
      }
      // A_next = Pb[t] + Pnb[t]  (Counter addition: keys of Pb[t] first, then keys only in Pnb[t]; only positive sums survive)
      cand.clear();
      cand_val.clear();
      for (int32_t n : order_b) {
        const double v = pb_cur[n] + (snb_cur[n] == t ? pnb_cur[n] : 0.0);
        if (v > 0) {
          cand.push_back(n);
          cand_val.push_back(v);
        }
      }
      for (int32_t n : order_nb)
        if (sb_cur[n] != t && pnb_cur[n] > 0) {
          cand.push_back(n);
          cand_val.push_back(pnb_cur[n]);
        }
      cand_key.resize(cand.size());
      perm.resize(cand.size());
      for (size_t i = 0; i < cand.size(); ++i) {
        cand_key[i] = cand_val[i] * word_bonus(cand[i]);
        perm[i] = (int32_t)i;
      }
      std::stable_sort(perm.begin(), perm.end(), [&](int32_t a, int32_t b) { return cand_key[a] > cand_key[b]; });
      beam.clear();
      for (size_t i = 0; i < perm.size() && (int32_t)i < k; ++i) beam.push_back(cand[perm[i]]);
      if (score_out) *score_out = perm.empty() ? 0.0 : cand_key[perm[0]];
      pb_prev.swap(pb_cur);
      pnb_prev.swap(pnb_cur);
      sb_prev.swap(sb_cur);
      snb_prev.swap(snb_cur);
    }
    return beam.empty() ? 0 : beam[0];
  }
};

}  // namespace

extern "C" int w2l_prefix_beam_search_host(const double* probs_host, const int64_t* frames_host, int64_t n_utt, int64_t T_max, int64_t F,
                                           int32_t blank, int32_t space_id, int32_t end_id, const uint8_t* is_word_host,
                                           const uint8_t* is_term_host, int32_t k, double alpha, double beta, double prune,
                                           w2l_lm_callback lm, void* lm_user, int32_t* out_ids_host, int64_t out_stride,
                                           int64_t* out_len_host, double* out_score_host, int32_t threads) {
  if (!probs_host || !is_word_host || !is_term_host || !out_ids_host || !out_len_host || n_utt < 0 || T_max < 1 || F < 1 || blank < 0 ||
      blank >= F || k < 1)
    return W2L_ERR_INVALID_ARGUMENT;
  auto work = [&](int64_t u) {
    Search s;
    s.probs = probs_host + u * T_max * F;
    s.T = frames_host ? std::min<int64_t>(std::max<int64_t>(frames_host[u], 0), T_max) : T_max;
    s.F = F;
    s.blank = blank;
    s.space_id = space_id;
    s.end_id = end_id;
    s.k = k;
    s.alpha = alpha;
    s.beta = beta;
    s.prune = prune;
    s.is_word = is_word_host;
    s.is_term = is_term_host;
    s.lm = lm;
    s.lm_user = lm_user;
    double score = 0.0;
    const int32_t best = s.run(&score);
    std::vector<int32_t> ids;
    s.ids_of(best, ids);
    const int64_t n = std::min<int64_t>((int64_t)ids.size(), out_stride);
    std::memcpy(out_ids_host + u * out_stride, ids.data(), sizeof(int32_t) * (size_t)n);
    out_len_host[u] = (int64_t)ids.size();
    if (out_score_host) out_score_host[u] = score;
  };
  int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
  if (nt < 1) nt = 1;
  if (nt > 32) nt = 32;
  if (lm) nt = 1;                                        // a Python callback must not be entered from foreign threads
  if ((int64_t)nt > n_utt) nt = (int)n_utt;
  if (nt <= 1) {
    for (int64_t u = 0; u < n_utt; ++u) work(u);
    return W2L_OK;
  }
  std::vector<std::thread> pool;
  for (int i = 0; i < nt; ++i)
    pool.emplace_back([&, i]() {
      for (int64_t u = i; u < n_utt; u += nt) work(u);
    });
  for (auto& th : pool) th.join();
  return W2L_OK;
}
