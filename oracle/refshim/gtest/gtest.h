// gtest/gtest.h — minimal stand-in for the slice of GoogleTest that the reference's unit tests use.
//
// TEST INFRASTRUCTURE ONLY (see oracle/refshim/README.md).  GoogleTest is not installed in this image.  This
// header lets dpose_core/test/*.cpp of the reference compile UNMODIFIED: TEST / TEST_F / TEST_P,
// INSTANTIATE_TEST_SUITE_P with Values / ValuesIn / Range, EXPECT_* / ASSERT_* with streamed messages.
// Semantics follow GoogleTest: a fresh fixture object per test and parameter, GetParam() usable inside the
// fixture's constructor, Range(b, e, s) = {b, b+s, b+s+s, ...} while < e (repeated addition), ASSERT_* leaves the
// test body, EXPECT_* records and continues.  gtest_main.cpp next to this file provides main().
#ifndef DPOSE_REFSHIM_GTEST
#define DPOSE_REFSHIM_GTEST

#include <cstdio>
#include <functional>
#include <iterator>
#include <memory>
#include <tuple>
#include <exception>
#include <sstream>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

namespace testing {

class Test {
public:
  virtual ~Test() = default;
  virtual void SetUp() {}
  virtual void TearDown() {}
  virtual void TestBody() = 0;
};

template <typename T>
class WithParamInterface {
public:
  using ParamType = T;
  virtual ~WithParamInterface() = default;
  static const T& GetParam() { return *slot(); }
  static const T*& slot() {
    static const T* p = nullptr;
    return p;
  }
};

template <typename T>
class TestWithParam : public Test, public WithParamInterface<T> {};

class Message {
public:
  template <typename T>
  Message& operator<<(const T& v) {
    os_ << v;
    return *this;
  }
  std::string str() const { return os_.str(); }

private:
  std::ostringstream os_;
};

namespace internal {

struct State {
  int run = 0, failed = 0;
  bool current_failed = false;
  static State& get() {
    static State s;
    return s;
  }
};

struct Entry {
  std::string name;
  std::function<void()> run;
};

inline std::vector<Entry>& plain_tests() {
  static std::vector<Entry> v;
  return v;
}
/// TEST_P bodies of one fixture, expanded over its parameters when the tests run
template <typename Fixture>
std::vector<std::pair<std::string, std::function<void()>>>& param_tests() {
  static std::vector<std::pair<std::string, std::function<void()>>> v;
  return v;
}
inline std::vector<std::function<void(std::vector<Entry>&)>>& instantiations() {
  static std::vector<std::function<void(std::vector<Entry>&)>> v;
  return v;
}

template <typename T>
void run_one() {
  T t;
  t.SetUp();
  t.TestBody();
  t.TearDown();
}

struct Registrar {
  Registrar(const char* suite, const char* name, std::function<void()> fn) {
    plain_tests().push_back({std::string(suite) + "." + name, std::move(fn)});
  }
};
template <typename Fixture>
struct ParamRegistrar {
  ParamRegistrar(const char* suite, const char* name, std::function<void()> fn) {
    param_tests<Fixture>().emplace_back(std::string(suite) + "." + name, std::move(fn));
  }
};
template <typename Fixture>
struct Instantiator {
  template <typename Gen>
  Instantiator(const char* prefix, Gen gen) {
    using P = typename Fixture::ParamType;
    const std::string pre = prefix;
    instantiations().push_back([pre, gen](std::vector<Entry>& out) {
      const auto params = std::make_shared<std::vector<P>>(gen.template get<P>());
      for (const auto& t : param_tests<Fixture>())
        for (size_t i = 0; i < params->size(); ++i) {
          auto body = t.second;
          out.push_back({(pre.empty() ? "" : pre + "/") + t.first + "/" + std::to_string(i), [params, i, body] {
                           WithParamInterface<P>::slot() = &(*params)[i];
                           body();
                           WithParamInterface<P>::slot() = nullptr;
                         }});
        }
    });
  }
};

// ---- value printing: streamed when the type has operator<<, a placeholder otherwise
template <typename T, typename = void>
struct streamable : std::false_type {};
template <typename T>
struct streamable<T, decltype(void(std::declval<std::ostream&>() << std::declval<const T&>()))> : std::true_type {};
template <typename T>
std::string show(const T& v, std::true_type) {
  std::ostringstream os;
  os << v;
  return os.str();
}
template <typename T>
std::string show(const T&, std::false_type) {
  return "<unprintable>";
}
template <typename T>
std::string show(const T& v) {
  return show(v, streamable<T>());
}

struct Result {
  bool ok;
  std::string text;
  explicit operator bool() const { return ok; }
};

struct Failure {
  Failure(const char* file, int line, std::string text) : file_(file), line_(line), text_(std::move(text)) {}
  void operator=(const Message& m) const {
    State::get().current_failed = true;
    std::printf("%s:%d: Failure\n%s\n%s\n", file_, line_, text_.c_str(), m.str().c_str());
  }
  const char* file_;
  int line_;
  std::string text_;
};

#define DPOSE_GTEST_CMP_(name, op)                                                                     \
  template <typename A, typename B>                                                                    \
  Result name(const char* ea, const char* eb, const A& a, const B& b) {                                \
    if (a op b) return {true, ""};                                                                     \
    return {false, std::string("Expected: (") + ea + ") " #op " (" + eb + "), actual: " + show(a) + " vs " + show(b)}; \
  }
DPOSE_GTEST_CMP_(cmp_eq, ==)
DPOSE_GTEST_CMP_(cmp_ne, !=)
DPOSE_GTEST_CMP_(cmp_lt, <)
DPOSE_GTEST_CMP_(cmp_le, <=)
DPOSE_GTEST_CMP_(cmp_gt, >)
DPOSE_GTEST_CMP_(cmp_ge, >=)
#undef DPOSE_GTEST_CMP_

template <typename A, typename B, typename T>
Result near(const char* ea, const char* eb, const char* et, const A& a, const B& b, const T& tol) {
  const double diff = a > b ? (double)(a - b) : (double)(b - a);
  if (diff <= (double)tol) return {true, ""};
  return {false, std::string("The difference between ") + ea + " and " + eb + " is " + show(diff) + ", which exceeds " + et +
                     " (" + show(a) + " vs " + show(b) + ")"};
}

inline Result truth(const char* e, bool v, bool want) {
  if (v == want) return {true, ""};
  return {false, std::string("Value of: ") + e + "\n  Actual: " + (v ? "true" : "false") + "\nExpected: " + (want ? "true" : "false")};
}

// ---- parameter generators
template <typename... Ts>
struct ValuesGen {
  std::tuple<Ts...> vals;
  template <typename P>
  std::vector<P> get() const {
    return get_impl<P>(std::index_sequence_for<Ts...>());
  }
  template <typename P, size_t... I>
  std::vector<P> get_impl(std::index_sequence<I...>) const {
    return std::vector<P>{static_cast<P>(std::get<I>(vals))...};
  }
};
template <typename T>
struct VectorGen {
  std::vector<T> vals;
  template <typename P>
  std::vector<P> get() const {
    return std::vector<P>(vals.begin(), vals.end());
  }
};

}  // namespace internal

template <typename... Ts>
internal::ValuesGen<Ts...> Values(Ts... vs) {
  return {std::make_tuple(vs...)};
}
template <typename T, size_t N>
internal::VectorGen<T> ValuesIn(const T (&a)[N]) {
  return {std::vector<T>(a, a + N)};
}
template <typename C>
internal::VectorGen<typename C::value_type> ValuesIn(const C& c) {
  return {std::vector<typename C::value_type>(std::begin(c), std::end(c))};
}
template <typename T, typename S>
internal::VectorGen<T> Range(T begin, T end, S step) {
  internal::VectorGen<T> g;
  for (T i = begin; i < end; i = static_cast<T>(i + step)) g.vals.push_back(i);
  return g;
}
template <typename T>
internal::VectorGen<T> Range(T begin, T end) {
  return Range(begin, end, T(1));
}

inline void InitGoogleTest(int*, char**) {}

inline int RunAllTests() {
  using namespace internal;
  std::vector<Entry> all = plain_tests();
  for (const auto& inst : instantiations()) inst(all);
  State& s = State::get();
  for (const auto& t : all) {
    s.current_failed = false;
    try {
      t.run();
    } catch (const std::exception& e) {
      s.current_failed = true;
      std::printf("unexpected exception in test body: %s\n", e.what());
    } catch (...) {
      s.current_failed = true;
      std::printf("unexpected exception in test body\n");
    }
    ++s.run;
    if (s.current_failed) {
      ++s.failed;
      std::printf("[  FAILED  ] %s\n", t.name.c_str());
    }
  }
  std::printf("[==========] %d tests ran. [  PASSED  ] %d. [  FAILED  ] %d.\n", s.run, s.run - s.failed, s.failed);
  return s.failed ? 1 : 0;
}

}  // namespace testing

#define RUN_ALL_TESTS() ::testing::RunAllTests()

#define DPOSE_GTEST_CLASS_(suite, name) suite##_##name##_Test

#define TEST(suite, name)                                                                               \
  class DPOSE_GTEST_CLASS_(suite, name) : public ::testing::Test {                                       \
  public:                                                                                                \
    void TestBody() override;                                                                            \
  };                                                                                                     \
  static ::testing::internal::Registrar suite##_##name##_registrar(#suite, #name,                        \
                                                                   &::testing::internal::run_one<DPOSE_GTEST_CLASS_(suite, name)>); \
  void DPOSE_GTEST_CLASS_(suite, name)::TestBody()

#define TEST_F(fixture, name)                                                                           \
  class DPOSE_GTEST_CLASS_(fixture, name) : public fixture {                                             \
  public:                                                                                                \
    void TestBody() override;                                                                            \
  };                                                                                                     \
  static ::testing::internal::Registrar fixture##_##name##_registrar(#fixture, #name,                    \
                                                                     &::testing::internal::run_one<DPOSE_GTEST_CLASS_(fixture, name)>); \
  void DPOSE_GTEST_CLASS_(fixture, name)::TestBody()

#define TEST_P(fixture, name)                                                                           \
  class DPOSE_GTEST_CLASS_(fixture, name) : public fixture {                                             \
  public:                                                                                                \
    void TestBody() override;                                                                            \
  };                                                                                                     \
  static ::testing::internal::ParamRegistrar<fixture> fixture##_##name##_registrar(                      \
      #fixture, #name, &::testing::internal::run_one<DPOSE_GTEST_CLASS_(fixture, name)>);                \
  void DPOSE_GTEST_CLASS_(fixture, name)::TestBody()

#define INSTANTIATE_TEST_SUITE_P(prefix, fixture, ...) \
  static ::testing::internal::Instantiator<fixture> prefix##fixture##_instantiation(#prefix, __VA_ARGS__)
#define INSTANTIATE_TEST_CASE_P(prefix, fixture, ...) INSTANTIATE_TEST_SUITE_P(prefix, fixture, __VA_ARGS__)

// `switch (0) case 0: default:` swallows a dangling else, as in GoogleTest
#define DPOSE_GTEST_CHECK_(result, on_failure)                   \
  switch (0)                                                     \
  case 0:                                                        \
  default:                                                       \
    if (const ::testing::internal::Result gt_res_ = (result))    \
      ;                                                          \
    else                                                         \
      on_failure ::testing::internal::Failure(__FILE__, __LINE__, gt_res_.text) = ::testing::Message()

#define DPOSE_GTEST_NONFATAL_
#define DPOSE_GTEST_FATAL_ return

#define EXPECT_EQ(a, b) DPOSE_GTEST_CHECK_(::testing::internal::cmp_eq(#a, #b, a, b), DPOSE_GTEST_NONFATAL_)
#define EXPECT_NE(a, b) DPOSE_GTEST_CHECK_(::testing::internal::cmp_ne(#a, #b, a, b), DPOSE_GTEST_NONFATAL_)
#define EXPECT_LT(a, b) DPOSE_GTEST_CHECK_(::testing::internal::cmp_lt(#a, #b, a, b), DPOSE_GTEST_NONFATAL_)
#define EXPECT_LE(a, b) DPOSE_GTEST_CHECK_(::testing::internal::cmp_le(#a, #b, a, b), DPOSE_GTEST_NONFATAL_)
#define EXPECT_GT(a, b) DPOSE_GTEST_CHECK_(::testing::internal::cmp_gt(#a, #b, a, b), DPOSE_GTEST_NONFATAL_)
#define EXPECT_GE(a, b) DPOSE_GTEST_CHECK_(::testing::internal::cmp_ge(#a, #b, a, b), DPOSE_GTEST_NONFATAL_)
#define EXPECT_NEAR(a, b, tol) DPOSE_GTEST_CHECK_(::testing::internal::near(#a, #b, #tol, a, b, tol), DPOSE_GTEST_NONFATAL_)
#define ASSERT_NEAR(a, b, tol) DPOSE_GTEST_CHECK_(::testing::internal::near(#a, #b, #tol, a, b, tol), DPOSE_GTEST_FATAL_)
#define EXPECT_TRUE(c) DPOSE_GTEST_CHECK_(::testing::internal::truth(#c, static_cast<bool>(c), true), DPOSE_GTEST_NONFATAL_)
#define EXPECT_FALSE(c) DPOSE_GTEST_CHECK_(::testing::internal::truth(#c, static_cast<bool>(c), false), DPOSE_GTEST_NONFATAL_)
#define ASSERT_EQ(a, b) DPOSE_GTEST_CHECK_(::testing::internal::cmp_eq(#a, #b, a, b), DPOSE_GTEST_FATAL_)
#define ASSERT_NE(a, b) DPOSE_GTEST_CHECK_(::testing::internal::cmp_ne(#a, #b, a, b), DPOSE_GTEST_FATAL_)
#define ASSERT_LT(a, b) DPOSE_GTEST_CHECK_(::testing::internal::cmp_lt(#a, #b, a, b), DPOSE_GTEST_FATAL_)
#define ASSERT_LE(a, b) DPOSE_GTEST_CHECK_(::testing::internal::cmp_le(#a, #b, a, b), DPOSE_GTEST_FATAL_)
#define ASSERT_GT(a, b) DPOSE_GTEST_CHECK_(::testing::internal::cmp_gt(#a, #b, a, b), DPOSE_GTEST_FATAL_)
#define ASSERT_GE(a, b) DPOSE_GTEST_CHECK_(::testing::internal::cmp_ge(#a, #b, a, b), DPOSE_GTEST_FATAL_)
#define ASSERT_TRUE(c) DPOSE_GTEST_CHECK_(::testing::internal::truth(#c, static_cast<bool>(c), true), DPOSE_GTEST_FATAL_)
#define ASSERT_FALSE(c) DPOSE_GTEST_CHECK_(::testing::internal::truth(#c, static_cast<bool>(c), false), DPOSE_GTEST_FATAL_)

#define DPOSE_GTEST_THROWS_(stmt, on_failure)                                                          \
  DPOSE_GTEST_CHECK_(([&]() -> ::testing::internal::Result {                                            \
                       try {                                                                            \
                         stmt;                                                                          \
                       } catch (...) {                                                                  \
                         return {true, ""};                                                            \
                       }                                                                                \
                       return {false, "Expected: " #stmt " throws an exception.\n  Actual: it doesn't."}; \
                     })(),                                                                              \
                     on_failure)
#define EXPECT_ANY_THROW(stmt) DPOSE_GTEST_THROWS_(stmt, DPOSE_GTEST_NONFATAL_)
#define ASSERT_ANY_THROW(stmt) DPOSE_GTEST_THROWS_(stmt, DPOSE_GTEST_FATAL_)

#endif  // DPOSE_REFSHIM_GTEST
