// ORACLE — TEST INFRASTRUCTURE ONLY (see pvo_math.hpp header).  Parity status: unpinned by the reference's own tests
// (it has none); pinned here against scipy connected components (tests/test_builders.py).
//
// LiDAR line tracks: util/Tracks.h:34-110 (UnionFind), util/Tracks.cpp:58-186 (TrackBuilder::Build / Filter / ExportTracks)
// as driven by lidar_mapping/LidarLineMatch.cpp:36-86 (GenerateTracks) and consumed by util/Optimization.cpp:343-400.
#pragma once
#include <cstdint>
#include <limits>
#include <map>
#include <numeric>
#include <set>
#include <utility>
#include <vector>

namespace pvo {

typedef std::pair<uint32_t, uint32_t> Feature;   // {frame id, line id}

struct DisjointSets {   // union by rank + path compression, with a per-root size (Tracks.h:34-110)
  std::vector<uint32_t> parent, rank, size;
  void Init(uint32_t n) { parent.resize(n); std::iota(parent.begin(), parent.end(), 0u); rank.assign(n, 0u); size.assign(n, 1u); }
  uint32_t Find(uint32_t i) { if (parent[i] != i) parent[i] = Find(parent[i]); return parent[i]; }
  void Union(uint32_t i, uint32_t j) {
    i = Find(i); j = Find(j);
    if (i == j) return;
    if (rank[i] < rank[j]) { parent[i] = j; size[j] += size[i]; }
    else { parent[j] = i; size[i] += size[j]; if (rank[i] == rank[j]) ++rank[i]; }
  }
};

struct LineTracks { std::vector<std::set<Feature>> tracks; };   // tracks[t] = LineTrack::feature_pairs, id = t (LidarLineMatch.cpp:80-81)

// pairs[p] = {frame a, frame b}; matches[p] = set of {line of a, line of b}  (Tracks.cpp:58-101)
inline void BuildLineTracks(const std::vector<std::pair<size_t, size_t>>& pairs, const std::vector<std::set<Feature>>& matches, uint32_t min_length,
                            bool allow_multiple_map, LineTracks& out) {
  std::set<Feature> all;
  for (size_t p = 0; p < pairs.size(); ++p)
    for (const Feature& m : matches[p]) { all.emplace((uint32_t)pairs[p].first, m.first); all.emplace((uint32_t)pairs[p].second, m.second); }
  std::map<Feature, uint32_t> f2i; std::map<uint32_t, Feature> i2f;
  uint32_t count = 0;
  for (const Feature& f : all) { f2i.emplace(f, count); i2f.emplace(count, f); ++count; }
  DisjointSets uf; uf.Init((uint32_t)f2i.size());
  for (size_t p = 0; p < pairs.size(); ++p)
    for (const Feature& m : matches[p]) uf.Union(f2i[Feature((uint32_t)pairs[p].first, m.first)], f2i[Feature((uint32_t)pairs[p].second, m.second)]);
  // Filter (Tracks.cpp:103-139)
  const uint32_t kBad = std::numeric_limits<uint32_t>::max();
  std::map<uint32_t, std::set<uint32_t>> frames_of; std::set<uint32_t> bad;
  for (uint32_t i = 0; i < f2i.size(); ++i) {
    const uint32_t t = uf.Find(i);
    if (!frames_of[t].insert(i2f[i].first).second && !allow_multiple_map) bad.insert(t);
  }
  for (auto& kv : frames_of) if (kv.second.size() < min_length) bad.insert(kv.first);
  for (uint32_t& r : uf.parent) if (bad.count(r)) { uf.size[r] = 1; r = kBad; }
  // ExportTracks(std::vector<LineTrack>&) (Tracks.cpp:160-186)
  std::map<uint32_t, size_t> where;
  out.tracks.clear();
  for (uint32_t i = 0; i < f2i.size(); ++i) {
    const uint32_t t = uf.parent[i];
    if (t == kBad || !(uf.size[t] > 1)) continue;
    auto it = where.find(t);
    if (it != where.end()) out.tracks[it->second].insert(i2f[i]);
    else { where[t] = out.tracks.size(); out.tracks.push_back(std::set<Feature>{i2f[i]}); }
  }
}

// the gate of AddLidarLineToLineResidual2 (Optimization.cpp:343-400): keep an association only if the reference line belongs
// to a track that also contains the neighbour line
inline bool TrackGate(const LineTracks& T, const std::map<Feature, std::vector<uint32_t>>& lines_to_track, const Feature& ref, const Feature& nei) {
  auto it = lines_to_track.find(ref);
  if (it == lines_to_track.end()) return false;
  for (uint32_t t : it->second) if (T.tracks[t].count(nei) > 0) return true;
  return false;
}
inline std::map<Feature, std::vector<uint32_t>> LinesToTrack(const LineTracks& T) {
  std::map<Feature, std::vector<uint32_t>> m;
  for (uint32_t t = 0; t < T.tracks.size(); ++t) for (const Feature& f : T.tracks[t]) m[f].push_back(t);
  return m;
}

}  // namespace pvo
