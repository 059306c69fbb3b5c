// ByteTrack (Zhang et al., ECCV 2022) on the host, in C++: the sequential, order-dependent stage that follows the
// players detector (/root/reference/trackers/players_tracker/players_tracker.py:311,367-369 -- sv.ByteTrack(frame_rate)
// .update_with_detections).  It runs on rank 0 for every frame of every shard, after the gather; at several thousand
// frames per second the numpy version (trackers/sv_compat.py::ByteTrack, ~0.3 ms per frame) becomes the slowest
// stage of the whole pass, this one costs a few microseconds per frame.  Same algorithm, same constants, double
// precision like numpy; tests/test_host_cpu.py checks id-for-id equality with the Python version.
//
// No CUDA in this file (host code of the shared library).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

#include "internal.h"

namespace pb {
namespace {

constexpr double kWPos = 1.0 / 20, kWVel = 1.0 / 160;
enum { ST_NEW = 0, ST_TRACKED = 1, ST_LOST = 2, ST_REMOVED = 3 };

struct Trk {
  double mean[8];
  double cov[64];
  double tlbr0[4];
  double score;
  int state = ST_NEW, id = 0, frame_id = 0, start_frame = 0, tracklet_len = 0, det_index = -1;
  bool activated = false, has_mean = false;
};

inline void tlbr_of(const Trk& t, double* o) {
  if (!t.has_mean) {
    memcpy(o, t.tlbr0, sizeof(double) * 4);
    return;
  }
  const double x = t.mean[0], y = t.mean[1], a = t.mean[2], h = t.mean[3], w = a * h;
  o[0] = x - w / 2;
  o[1] = y - h / 2;
  o[2] = x + w / 2;
  o[3] = y + h / 2;
}

inline double iou(const double* a, const double* b) {
  const double x1 = std::max(a[0], b[0]), y1 = std::max(a[1], b[1]);
  const double x2 = std::min(a[2], b[2]), y2 = std::min(a[3], b[3]);
  const double inter = std::max(x2 - x1, 0.0) * std::max(y2 - y1, 0.0);
  const double uni = (a[2] - a[0]) * (a[3] - a[1]) + (b[2] - b[0]) * (b[3] - b[1]) - inter;
  return uni > 0 ? inter / uni : 0.0;
}

void kf_initiate(Trk& t) {
  const double* b = t.tlbr0;
  const double w = b[2] - b[0], h = b[3] - b[1];
  const double m[4] = {b[0] + w / 2, b[1] + h / 2, w / h, h};
  for (int i = 0; i < 4; ++i) t.mean[i] = m[i], t.mean[4 + i] = 0;
  const double std[8] = {2 * kWPos * h, 2 * kWPos * h, 1e-2, 2 * kWPos * h, 10 * kWVel * h, 10 * kWVel * h, 1e-5, 10 * kWVel * h};
  memset(t.cov, 0, sizeof(t.cov));
  for (int i = 0; i < 8; ++i) t.cov[i * 9] = std[i] * std[i];
  t.has_mean = true;
}

// x' = F x, P' = F P F^T + Q with F = [[I, I], [0, I]]
void kf_predict(Trk& t) {
  if (t.state != ST_TRACKED) t.mean[7] = 0;
  const double h = t.mean[3];
  const double std[8] = {kWPos * h, kWPos * h, 1e-2, kWPos * h, kWVel * h, kWVel * h, 1e-5, kWVel * h};
  for (int i = 0; i < 4; ++i) t.mean[i] += t.mean[4 + i];
  double fp[64], out[64];
  for (int i = 0; i < 8; ++i)  // F P: rows 0..3 get row i + row i+4
    for (int j = 0; j < 8; ++j) fp[i * 8 + j] = t.cov[i * 8 + j] + (i < 4 ? t.cov[(i + 4) * 8 + j] : 0.0);
  for (int i = 0; i < 8; ++i)  // (F P) F^T: columns 0..3 get col j + col j+4
    for (int j = 0; j < 8; ++j) out[i * 8 + j] = fp[i * 8 + j] + (j < 4 ? fp[i * 8 + j + 4] : 0.0);
  for (int i = 0; i < 8; ++i) out[i * 9] += std[i] * std[i];
  memcpy(t.cov, out, sizeof(out));
}

// measurement z = (x, y, a, h); H selects the first four states
void kf_update(Trk& t, const double* z) {
  const double h = t.mean[3];
  const double std[4] = {kWPos * h, kWPos * h, 1e-1, kWPos * h};
  double S[16], Sinv[16];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) S[i * 4 + j] = t.cov[i * 8 + j] + (i == j ? std[i] * std[i] : 0.0);
  {  // 4x4 inverse by Gauss-Jordan with partial pivoting (S is symmetric positive definite)
    double a[4][8];
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) a[i][j] = S[i * 4 + j], a[i][4 + j] = i == j ? 1.0 : 0.0;
    for (int c = 0; c < 4; ++c) {
      int p = c;
      for (int r = c + 1; r < 4; ++r)
        if (std::fabs(a[r][c]) > std::fabs(a[p][c])) p = r;
      if (p != c)
        for (int j = 0; j < 8; ++j) std::swap(a[p][j], a[c][j]);
      const double d = a[c][c];
      for (int j = 0; j < 8; ++j) a[c][j] /= d;
      for (int r = 0; r < 4; ++r)
        if (r != c) {
          const double f = a[r][c];
          for (int j = 0; j < 8; ++j) a[r][j] -= f * a[c][j];
        }
    }
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) Sinv[i * 4 + j] = a[i][4 + j];
  }
  double K[32];  // (8,4) = P[:, :4] S^-1
  for (int i = 0; i < 8; ++i)
    for (int j = 0; j < 4; ++j) {
      double s = 0;
      for (int k = 0; k < 4; ++k) s += t.cov[i * 8 + k] * Sinv[k * 4 + j];
      K[i * 4 + j] = s;
    }
  double innov[4];
  for (int i = 0; i < 4; ++i) innov[i] = z[i] - t.mean[i];
  for (int i = 0; i < 8; ++i)
    for (int k = 0; k < 4; ++k) t.mean[i] += K[i * 4 + k] * innov[k];
  double KS[32];
  for (int i = 0; i < 8; ++i)
    for (int j = 0; j < 4; ++j) {
      double s = 0;
      for (int k = 0; k < 4; ++k) s += K[i * 4 + k] * S[k * 4 + j];
      KS[i * 4 + j] = s;
    }
  for (int i = 0; i < 8; ++i)
    for (int j = 0; j < 8; ++j) {
      double s = 0;
      for (int k = 0; k < 4; ++k) s += KS[i * 4 + k] * K[j * 4 + k];
      t.cov[i * 8 + j] -= s;
    }
}

// Minimum-cost assignment of an R x C matrix (Hungarian algorithm with potentials, O(n^3)); entries > thresh are
// capped at thresh + 1e-4 first and pairs above thresh are not reported (supervision matching.linear_assignment).
void linear_assignment(std::vector<double> cost, int R, int C, double thresh, std::vector<std::pair<int, int>>& matches,
                       std::vector<int>& un_r, std::vector<int>& un_c) {
  matches.clear();
  un_r.clear();
  un_c.clear();
  if (R == 0 || C == 0) {
    for (int i = 0; i < R; ++i) un_r.push_back(i);
    for (int j = 0; j < C; ++j) un_c.push_back(j);
    return;
  }
  for (auto& v : cost)
    if (v > thresh) v = thresh + 1e-4;
  const bool tr = R > C;  // the algorithm wants rows <= columns
  const int n = tr ? C : R, m = tr ? R : C;
  auto at = [&](int i, int j) { return tr ? cost[(size_t)j * C + i] : cost[(size_t)i * C + j]; };
  const double INF = std::numeric_limits<double>::infinity();
  std::vector<double> u(n + 1, 0.0), v(m + 1, 0.0), minv(m + 1);
  std::vector<int> p(m + 1, 0), way(m + 1, 0);
  std::vector<char> used(m + 1);
  for (int i = 1; i <= n; ++i) {
    p[0] = i;
    int j0 = 0;
    std::fill(minv.begin(), minv.end(), INF);
    std::fill(used.begin(), used.end(), 0);
    do {
      used[j0] = 1;
      const int i0 = p[j0];
      double delta = INF;
      int j1 = 0;
      for (int j = 1; j <= m; ++j)
        if (!used[j]) {
          const double cur = at(i0 - 1, j - 1) - u[i0] - v[j];
          if (cur < minv[j]) minv[j] = cur, way[j] = j0;
          if (minv[j] < delta) delta = minv[j], j1 = j;
        }
      for (int j = 0; j <= m; ++j)
        if (used[j]) u[p[j]] += delta, v[j] -= delta;
        else minv[j] -= delta;
      j0 = j1;
    } while (p[j0] != 0);
    do {
      const int j1 = way[j0];
      p[j0] = p[j1];
      j0 = j1;
    } while (j0);
  }
  std::vector<char> mr(R, 0), mc(C, 0);
  for (int j = 1; j <= m; ++j)
    if (p[j]) {
      const int r = tr ? j - 1 : p[j] - 1, c = tr ? p[j] - 1 : j - 1;
      if (cost[(size_t)r * C + c] <= thresh) {
        matches.emplace_back(r, c);
        mr[r] = mc[c] = 1;
      }
    }
  std::sort(matches.begin(), matches.end());
  for (int i = 0; i < R; ++i)
    if (!mr[i]) un_r.push_back(i);
  for (int j = 0; j < C; ++j)
    if (!mc[j]) un_c.push_back(j);
}

}  // namespace
}  // namespace pb

struct pb_bytetrack {
  double act_thresh, match_thresh, det_thresh;
  int max_time_lost, frame_id = 0, next_id = 0;
  std::vector<pb::Trk*> tracked, lost;
  ~pb_bytetrack() { clear(); }
  void clear() {
    for (auto* t : tracked) delete t;
    for (auto* t : lost) delete t;
    tracked.clear();
    lost.clear();
    frame_id = next_id = 0;
  }
};

using namespace pb;

namespace {

std::vector<double> iou_cost(const std::vector<Trk*>& a, const std::vector<Trk*>& b, bool fuse) {
  std::vector<double> c(a.size() * b.size());
  std::vector<double> ba(a.size() * 4), bb(b.size() * 4);
  for (size_t i = 0; i < a.size(); ++i) tlbr_of(*a[i], &ba[i * 4]);
  for (size_t j = 0; j < b.size(); ++j) tlbr_of(*b[j], &bb[j * 4]);
  for (size_t i = 0; i < a.size(); ++i)
    for (size_t j = 0; j < b.size(); ++j) {
      const double sim = iou(&ba[i * 4], &bb[j * 4]);
      c[i * b.size() + j] = fuse ? 1.0 - sim * b[j]->score : 1.0 - sim;
    }
  return c;
}

void apply_matches(pb_bytetrack* bt, const std::vector<Trk*>& tracks, const std::vector<Trk*>& dets,
                   const std::vector<std::pair<int, int>>& matches, std::vector<Trk*>& activated,
                   std::vector<Trk*>& refind) {
  for (const auto& mt : matches) {
    Trk* t = tracks[mt.first];
    const Trk* d = dets[mt.second];
    const double* b = d->tlbr0;
    const double w = b[2] - b[0], h = b[3] - b[1];
    const double z[4] = {b[0] + w / 2, b[1] + h / 2, w / h, h};
    kf_update(*t, z);
    if (t->state == ST_TRACKED) {
      t->tracklet_len += 1;
      activated.push_back(t);
    } else {
      t->tracklet_len = 0;
      refind.push_back(t);
    }
    t->state = ST_TRACKED;
    t->activated = true;
    t->frame_id = bt->frame_id;
    t->score = d->score;
  }
}

bool contains(const std::vector<Trk*>& v, const Trk* t) { return std::find(v.begin(), v.end(), t) != v.end(); }

}  // namespace

extern "C" {

pb_bytetrack* pb_bytetrack_create(double track_activation_threshold, int lost_track_buffer,
                                  double minimum_matching_threshold, double frame_rate) {
  auto* bt = new pb_bytetrack();
  bt->act_thresh = track_activation_threshold;
  bt->match_thresh = minimum_matching_threshold;
  bt->det_thresh = track_activation_threshold + 0.1;
  bt->max_time_lost = (int)(frame_rate / 30.0 * lost_track_buffer);
  return bt;
}

void pb_bytetrack_destroy(pb_bytetrack* bt) { delete bt; }
void pb_bytetrack_reset(pb_bytetrack* bt) {
  if (bt) bt->clear();
}

/* boxes float (n,4) xyxy, scores float (n): one frame, in order.  ids_out int (n): track id of each detection, -1 for
 * detections without an active track (the caller drops those, as update_with_detections does). */
int pb_bytetrack_update(pb_bytetrack* bt, const float* boxes, const float* scores, int n, int* ids_out) {
  PB_CHECK(bt != nullptr && (n == 0 || (boxes && scores && ids_out)), "bytetrack_update: null argument");
  bt->frame_id += 1;
  std::vector<Trk*> dets, dets2, owned;
  for (int i = 0; i < n; ++i) {
    const double s = scores[i];
    const bool hi = s > bt->act_thresh, lo = s > 0.1 && s < bt->act_thresh;
    if (!hi && !lo) continue;
    auto* d = new Trk();
    for (int k = 0; k < 4; ++k) d->tlbr0[k] = boxes[i * 4 + k];
    d->score = s;
    d->det_index = i;
    owned.push_back(d);
    (hi ? dets : dets2).push_back(d);
  }
  std::vector<Trk*> unconfirmed, tracked, pool, activated, refind, lost_now;
  for (auto* t : bt->tracked) (t->activated ? tracked : unconfirmed).push_back(t);
  pool = tracked;
  for (auto* t : bt->lost)
    if (!contains(pool, t)) pool.push_back(t);
  for (auto* t : pool) kf_predict(*t);
  std::vector<std::pair<int, int>> matches;
  std::vector<int> u_track, u_det, u_track2, u_det2, u_unc;
  linear_assignment(iou_cost(pool, dets, true), (int)pool.size(), (int)dets.size(), bt->match_thresh, matches, u_track, u_det);
  apply_matches(bt, pool, dets, matches, activated, refind);
  std::vector<Trk*> r_tracked;
  for (int i : u_track)
    if (pool[i]->state == ST_TRACKED) r_tracked.push_back(pool[i]);
  linear_assignment(iou_cost(r_tracked, dets2, false), (int)r_tracked.size(), (int)dets2.size(), 0.5, matches, u_track2, u_det2);
  apply_matches(bt, r_tracked, dets2, matches, activated, refind);
  for (int i : u_track2)
    if (r_tracked[i]->state != ST_LOST) {
      r_tracked[i]->state = ST_LOST;
      lost_now.push_back(r_tracked[i]);
    }
  std::vector<Trk*> rest;
  for (int j : u_det) rest.push_back(dets[j]);
  linear_assignment(iou_cost(unconfirmed, rest, true), (int)unconfirmed.size(), (int)rest.size(), 0.7, matches, u_unc, u_det);
  apply_matches(bt, unconfirmed, rest, matches, activated, activated);
  for (int i : u_unc) unconfirmed[i]->state = ST_REMOVED;
  for (int j : u_det) {
    Trk* d = rest[j];
    if (d->score < bt->det_thresh) continue;
    auto* t = new Trk(*d);  // the detection becomes a track
    t->id = ++bt->next_id;
    kf_initiate(*t);
    t->tracklet_len = 0;
    t->state = ST_TRACKED;
    t->activated = bt->frame_id == 1;
    t->frame_id = t->start_frame = bt->frame_id;
    activated.push_back(t);
  }
  for (auto* t : bt->lost)
    if (bt->frame_id - t->frame_id > bt->max_time_lost) t->state = ST_REMOVED;
  // new tracked list: still-tracked old ones, then newly activated, then re-found (each once, by pointer)
  std::vector<Trk*> nt;
  for (auto* t : bt->tracked)
    if (t->state == ST_TRACKED) nt.push_back(t);
  for (auto* t : activated)
    if (!contains(nt, t)) nt.push_back(t);
  for (auto* t : refind)
    if (!contains(nt, t)) nt.push_back(t);
  std::vector<Trk*> nl;
  for (auto* t : bt->lost)
    if (!contains(nt, t)) nl.push_back(t);
  for (auto* t : lost_now) nl.push_back(t);
  std::vector<Trk*> dead;
  {
    std::vector<Trk*> keep;
    for (auto* t : nl) (t->state == ST_REMOVED ? dead : keep).push_back(t);
    nl.swap(keep);
  }
  for (auto* t : bt->tracked)
    if (t->state == ST_REMOVED && !contains(dead, t)) dead.push_back(t);
  // duplicate pruning: of a tracked / lost pair with IoU > 0.85 the one with the shorter history goes
  if (!nt.empty() && !nl.empty()) {
    const std::vector<double> pd = iou_cost(nt, nl, false);
    std::vector<char> da(nt.size(), 0), db(nl.size(), 0);
    for (size_t p = 0; p < nt.size(); ++p)
      for (size_t q = 0; q < nl.size(); ++q)
        if (pd[p * nl.size() + q] < 0.15) {
          const int tp = nt[p]->frame_id - nt[p]->start_frame, tq = nl[q]->frame_id - nl[q]->start_frame;
          if (tp > tq) db[q] = 1; else da[p] = 1;
        }
    std::vector<Trk*> a2, b2;
    for (size_t p = 0; p < nt.size(); ++p) (da[p] ? dead : a2).push_back(nt[p]);
    for (size_t q = 0; q < nl.size(); ++q) (db[q] ? dead : b2).push_back(nl[q]);
    nt.swap(a2);
    nl.swap(b2);
  }
  bt->tracked.swap(nt);
  bt->lost.swap(nl);
  std::sort(dead.begin(), dead.end());
  dead.erase(std::unique(dead.begin(), dead.end()), dead.end());
  for (auto* t : dead) delete t;
  for (auto* d : owned) delete d;
  // attach ids: detections <-> active tracks by IoU (threshold 0.5), unmatched detections get -1
  std::vector<Trk*> out;
  for (auto* t : bt->tracked)
    if (t->activated) out.push_back(t);
  for (int i = 0; i < n; ++i) ids_out[i] = -1;
  if (!out.empty() && n > 0) {
    std::vector<double> c((size_t)n * out.size()), tb(out.size() * 4);
    for (size_t j = 0; j < out.size(); ++j) tlbr_of(*out[j], &tb[j * 4]);
    for (int i = 0; i < n; ++i) {
      const double b[4] = {boxes[i * 4], boxes[i * 4 + 1], boxes[i * 4 + 2], boxes[i * 4 + 3]};
      for (size_t j = 0; j < out.size(); ++j) c[(size_t)i * out.size() + j] = 1.0 - iou(b, &tb[j * 4]);
    }
    std::vector<int> ur, uc;
    linear_assignment(c, n, (int)out.size(), 0.5, matches, ur, uc);
    for (const auto& mt : matches) ids_out[mt.first] = out[mt.second]->id;
  }
  return 0;
}

/* `frames` consecutive frames in one call: counts int (frames) detections per frame, boxes / scores / ids_out the
 * frames' detections back to back.  Identical to calling pb_bytetrack_update once per frame. */
int pb_bytetrack_update_many(pb_bytetrack* bt, const float* boxes, const float* scores, const int* counts, int frames,
                             int* ids_out) {
  PB_CHECK(bt != nullptr && frames >= 0 && (frames == 0 || counts), "bytetrack_update_many: null argument");
  size_t at = 0;
  for (int f = 0; f < frames; ++f) {
    const int n = counts[f];
    PB_CHECK(n >= 0, "bytetrack_update_many: negative count at frame %d", f);
    const int rc = pb_bytetrack_update(bt, boxes + 4 * at, scores + at, n, ids_out + at);
    if (rc != 0) return rc;
    at += (size_t)n;
  }
  return 0;
}

}  // extern "C"
