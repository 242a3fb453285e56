// Source-compatibility header for geneticsCRE's src/util.h: src/wrapper.cpp brackets every level of the schedule with a
// `Timer timer(exec, path_length, total_paths)` object and calls Timer::print_header() once (src/wrapper.cpp:203,250-273).
// Upstream the class measures wall and CPU time but all of its output is commented out (src/util.h:16,23-24); here a
// line per level is printed when GCRE_TIMER=1 is set, otherwise the object is inert.
#ifndef GCRE_UTIL_H
#define GCRE_UTIL_H

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

class Timer {
  typedef std::chrono::steady_clock clock_type;

 public:
  Timer(const JoinExec& exec, int path_length, uint64_t total_paths)
      : exec_(exec), length_(path_length), paths_(total_paths), t0_(clock_type::now()) {}

  ~Timer() {
    if (!wanted()) return;
    const double ms = std::chrono::duration<double, std::milli>(clock_type::now() - t0_).count();
    std::printf("[gcre timer] level %d  method %d  patients<=%d  paths %llu  permutations %d  %.3f ms  (%s)\n", length_, static_cast<int>(exec_.method),
                exec_.width_ul * 64, static_cast<unsigned long long>(paths_), exec_.iters_requested, ms, gs_instr_label.c_str());
  }

  static void print_header() {
    if (wanted()) std::printf("[gcre timer] one line per level follows\n");
  }

 private:
  static bool wanted() {
    const char* flag = std::getenv("GCRE_TIMER");
    return flag != nullptr && std::strcmp(flag, "0") != 0 && flag[0] != '\0';
  }

  const JoinExec& exec_;
  const int length_;
  const uint64_t paths_;
  const clock_type::time_point t0_;
};

#endif  // GCRE_UTIL_H
