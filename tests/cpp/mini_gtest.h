// mini_gtest.h — the handful of gtest macros the reference's test style needs (gtest is not installed here).
#pragma once
#include <csignal>
#include <cstdio>
#include <execinfo.h>
#include <unistd.h>
#include <functional>
#include <string>
#include <vector>
namespace mini_gtest {
struct Case { std::string name; std::function<void()> fn; };
inline std::vector<Case>& cases() { static std::vector<Case> c; return c; }
inline int& failures() { static int f = 0; return f; }
struct Reg { Reg(const char* s, const char* n, std::function<void()> f) { cases().push_back({std::string(s) + "." + n, f}); } };
inline void on_fatal_signal(int sig)
{
  const char msg[] = "\n*** fatal signal, backtrace (resolve with addr2line -e <binary>):\n";
  if (write(2, msg, sizeof(msg) - 1) < 0) {}
  void* frames[64];
  backtrace_symbols_fd(frames, backtrace(frames, 64), 2);
  _exit(128 + sig);
}
inline int run_all()
{
  std::setvbuf(stdout, nullptr, _IONBF, 0); // a crash must not swallow the lines before it
  std::signal(SIGSEGV, on_fatal_signal);
  std::signal(SIGABRT, on_fatal_signal);
  int bad = 0;
  for (auto& c : cases())
  {
    const int before = failures();
    std::printf("[ RUN      ] %s\n", c.name.c_str());
    c.fn();
    const bool ok = failures() == before;
    std::printf("[ %s ] %s\n", ok ? "      OK" : " FAILED ", c.name.c_str());
    bad += ok ? 0 : 1;
  }
  std::printf("%d test(s), %d failed\n", int(cases().size()), bad);
  return bad ? 1 : 0;
}
} // namespace mini_gtest
#define TEST(suite, name)                                                              \
  static void suite##_##name##_body();                                                 \
  static mini_gtest::Reg suite##_##name##_reg(#suite, #name, suite##_##name##_body);   \
  static void suite##_##name##_body()
#define EXPECT_EQ(a, b)                                                                                              \
  do { if (!((a) == (b))) { ++mini_gtest::failures(); std::printf("  %s:%d EXPECT_EQ(%s, %s) failed\n", __FILE__, __LINE__, #a, #b); } } while (0)
#define EXPECT_TRUE(a)                                                                                               \
  do { if (!(a)) { ++mini_gtest::failures(); std::printf("  %s:%d EXPECT_TRUE(%s) failed\n", __FILE__, __LINE__, #a); } } while (0)
#define EXPECT_FALSE(a)                                                                                              \
  do { if ((a)) { ++mini_gtest::failures(); std::printf("  %s:%d EXPECT_FALSE(%s) failed\n", __FILE__, __LINE__, #a); } } while (0)
#define RUN_ALL_TESTS() mini_gtest::run_all()
