// benchmark/benchmark.h — minimal stand-in for the slice of Google Benchmark that the reference's perf/*.cpp use.
//
// TEST INFRASTRUCTURE ONLY (see oracle/refshim/README.md).  Google Benchmark is not installed in this image.  This
// header lets dpose_core/perf/{dpose_core,dpose_costmap}.cpp compile UNMODIFIED: BENCHMARK(fn) with ->Arg / ->DenseRange,
// `for (auto _ : state)`, state.range(i), BENCHMARK_MAIN().  Like Google Benchmark it grows the iteration count
// until one run lasts at least the minimum time (0.5 s; DPOSE_BENCH_MIN_TIME overrides) and reports wall-clock
// nanoseconds per iteration of that run, as a table line and as a JSON line.
#ifndef DPOSE_REFSHIM_BENCHMARK
#define DPOSE_REFSHIM_BENCHMARK

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

namespace benchmark {

class State {
public:
  State(int64_t iterations, std::vector<int64_t> args) : iterations_(iterations), args_(std::move(args)) {}

  struct Value {};
  struct iterator {
    State* state;
    int64_t left;
    bool operator!=(const iterator&) {
      if (left != 0) return true;
      state->stop();
      return false;
    }
    iterator& operator++() {
      --left;
      return *this;
    }
    Value operator*() const { return Value(); }
  };
  iterator begin() {
    start_ = std::chrono::steady_clock::now();
    return iterator{this, iterations_};
  }
  iterator end() { return iterator{this, 0}; }

  int64_t range(size_t i = 0) const { return args_.at(i); }
  int64_t iterations() const { return iterations_; }
  double seconds() const { return seconds_; }
  void stop() { seconds_ = std::chrono::duration<double>(std::chrono::steady_clock::now() - start_).count(); }

private:
  int64_t iterations_;
  std::vector<int64_t> args_;
  std::chrono::steady_clock::time_point start_;
  double seconds_ = 0;
};

namespace internal {

struct Benchmark {
  std::string name;
  void (*fn)(State&);
  std::vector<std::vector<int64_t>> arg_sets;
  Benchmark* Arg(int64_t a) {
    arg_sets.push_back({a});
    return this;
  }
  Benchmark* DenseRange(int64_t lo, int64_t hi, int64_t step = 1) {
    for (int64_t a = lo; a <= hi; a += step) arg_sets.push_back({a});
    return this;
  }
};

inline std::vector<Benchmark*>& registry() {
  static std::vector<Benchmark*> v;
  return v;
}
inline Benchmark* add(const char* name, void (*fn)(State&)) {
  registry().push_back(new Benchmark{name, fn, {}});
  return registry().back();
}

inline int run_all() {
  const char* env = std::getenv("DPOSE_BENCH_MIN_TIME");
  const double min_time = env ? std::atof(env) : 0.5;
  std::printf("%-48s %15s %12s\n", "Benchmark", "Time", "Iterations");
  for (Benchmark* b : registry()) {
    std::vector<std::vector<int64_t>> sets = b->arg_sets;
    if (sets.empty()) sets.push_back({});
    for (const auto& args : sets) {
      std::string name = b->name;
      for (int64_t a : args) name += "/" + std::to_string(a);
      int64_t iters = 1;
      double secs = 0;
      for (;;) {
        State st(iters, args);
        b->fn(st);
        secs = st.seconds();
        if (secs >= min_time || iters >= (int64_t)1e9) break;
        const double grow = secs > 0 ? std::min(10.0, std::max(1.5, 1.4 * min_time / secs)) : 10.0;
        iters = (int64_t)(iters * grow) + 1;
      }
      const double ns = 1e9 * secs / (double)iters;
      std::printf("%-48s %12.1f ns %12lld\n", name.c_str(), ns, (long long)iters);
      std::printf("{\"benchmark\": \"%s\", \"real_time_ns\": %.3f, \"iterations\": %lld}\n", name.c_str(), ns, (long long)iters);
      std::fflush(stdout);
    }
  }
  return 0;
}

}  // namespace internal
}  // namespace benchmark

#define DPOSE_BENCH_CAT2_(a, b) a##b
#define DPOSE_BENCH_CAT_(a, b) DPOSE_BENCH_CAT2_(a, b)
#define BENCHMARK(fn) \
  static ::benchmark::internal::Benchmark* DPOSE_BENCH_CAT_(dpose_bench_, __LINE__) __attribute__((unused)) = ::benchmark::internal::add(#fn, fn)
#define BENCHMARK_MAIN() \
  int main(int, char**) { return ::benchmark::internal::run_all(); }

#endif  // DPOSE_REFSHIM_BENCHMARK
