// Drop-in replacement for the reference's src/util.h: the RAII Timer that src/wrapper.cpp wraps around each level
// (src/util.h:11-35; upstream all of its output is commented out).  Set GCRE_TIMER=1 to print one line per level.
#ifndef GCRE_UTIL_H
#define GCRE_UTIL_H

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <limits>
#include <unistd.h>
#include <vector>

class Timer {
 public:
  static void print_header() {
    if (enabled()) std::printf("\nTIME:PID IMPL METHOD WIDTH LENGTH PATHS PERMS MS\n\n");
  }

  Timer(const JoinExec& exec, int path_length, uint64_t total_paths) : exec_(exec), path_length_(path_length), total_paths_(total_paths) {}

  ~Timer() {
    if (!enabled()) return;
    const auto ms = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::system_clock::now() - start_).count();
    std::printf("\nTIME:%d %s m%d %d %d %lu %d %ld\n\n", (int)getpid(), gs_instr_label.c_str(), (int)exec_.method, exec_.width_ul * 64, path_length_,
                (unsigned long)total_paths_, exec_.iterations, (long)ms);
  }

 private:
  static bool enabled() {
    const char* e = std::getenv("GCRE_TIMER");
    return e && e[0] == '1';
  }
  const std::chrono::system_clock::time_point start_ = std::chrono::system_clock::now();
  const JoinExec& exec_;
  const int path_length_;
  const uint64_t total_paths_;
};

#endif
