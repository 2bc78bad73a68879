// word_orientation_study.cpp — CPU model of how many mask-word flushes (= RED instructions of raycast_dda_kernel) a scan
// costs for the three possible orientations of the update grid's 64-bit mask words: a word can hold the 8x8 voxels of an
// x-slice (today: OpenVDB's own layout, word = x&7), a y-slice or a z-slice of a leaf. A flush happens whenever the next
// voxel of a ray falls into another (leaf, word). Same DDA arithmetic as the oracle. Input: tools/bench_shim.py's file format.
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

static int32_t w2i(double c, double res)
{
  if (std::fmod(c, res)) c = c + (res / 2.0);
  return int32_t(std::floor(c * (1.0 / res)));
}

int main(int argc, char** argv)
{
  if (argc < 2) return 2;
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) return 2;
  double hdr[7];
  if (std::fread(hdr, sizeof(double), 7, f) != 7) return 2;
  const double res = hdr[0], range = hdr[1];
  double o[3];
  unsigned n = 0;
  if (std::fread(o, sizeof(double), 3, f) != 3 || std::fread(&n, sizeof(unsigned), 1, f) != 1) return 2;
  std::vector<float> pts(size_t(n) * 4);
  if (std::fread(pts.data(), 16, n, f) != n) return 2;
  std::fclose(f);
  const int32_t oi[3] = {w2i(o[0], res), w2i(o[1], res), w2i(o[2], res)};
  unsigned long long visits = 0, flush[3] = {0, 0, 0}, rays = 0;
  for (unsigned i = 0; i < n; ++i)
  {
    double e[3] = {pts[4 * i], pts[4 * i + 1], pts[4 * i + 2]};
    if (std::isnan(e[0]) || std::isnan(e[1]) || std::isnan(e[2])) continue;
    const double d[3] = {e[0] - o[0], e[1] - o[1], e[2] - o[2]};
    const double len  = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    if (len > range)
      for (int a = 0; a < 3; ++a) e[a] = o[a] + (d[a] / len) * range;
    const int32_t en[3] = {w2i(e[0], res), w2i(e[1], res), w2i(e[2], res)};
    if (en[0] == oi[0] && en[1] == oi[1] && en[2] == oi[2]) continue;
    ++rays;
    int32_t v[3] = {oi[0], oi[1], oi[2]}, step[3];
    double next[3], delta[3];
    for (int a = 0; a < 3; ++a)
    {
      const double dir = double(en[a]) - double(oi[a]);
      if (dir == 0.0) { step[a] = 0; next[a] = DBL_MAX; delta[a] = DBL_MAX; }
      else { delta[a] = std::fabs(1.0 / dir); next[a] = 0.5 * delta[a]; step[a] = dir > 0 ? 1 : -1; }
    }
    long long prev[3] = {-1, -1, -1};
    bool more;
    do
    {
      ++visits;
      const long long leaf = ((long long)(v[0] >> 3) * 2097152LL + (v[1] >> 3)) * 2097152LL + (v[2] >> 3);
      for (int a = 0; a < 3; ++a) // orientation a: the word index inside the leaf is coordinate a & 7
      {
        const long long w = leaf * 8 + (v[a] & 7);
        if (w != prev[a]) { ++flush[a]; prev[a] = w; }
      }
      const int axis = (next[0] < next[1] && next[0] < next[2]) ? 0 : ((next[1] < next[2]) ? 1 : 2);
      const double t = next[axis];
      next[axis] += delta[axis];
      v[axis] += step[axis];
      more = (t <= 1.0);
    } while (more);
  }
  std::printf("{\"rays\": %llu, \"visits\": %llu, \"flushes_x_slice_words\": %llu, \"flushes_y_slice_words\": %llu, \"flushes_z_slice_words\": %llu, "
              "\"visits_per_flush\": [%.2f, %.2f, %.2f]}\n",
              rays, visits, flush[0], flush[1], flush[2], double(visits) / flush[0], double(visits) / flush[1], double(visits) / flush[2]);
  return 0;
}
