// gtest.h -- a small stand-in for GoogleTest, TEST INFRASTRUCTURE ONLY.
//
// This image has no GoogleTest (and no network to fetch it), but the reference's current test suites
// (reference tests/ctest/*.cc) are written against it. This header implements the subset of the GoogleTest API those
// suites use, with GoogleTest's semantics, so that they compile UNMODIFIED from /root/reference against this library
// (oracle/ref_tests.mk):
//   TEST, TEST_F, TEST_P, INSTANTIATE_TEST_SUITE_P(prefix, suite, ValuesIn(...), name_fn), TestWithParam<T>,
//   TestParamInfo<T>, EXPECT_/ASSERT_{EQ,NE,LT,LE,GT,GE,TRUE,FALSE,STREQ,STRNE,THROW,NO_THROW} with streamed messages,
//   FAIL, ADD_FAILURE, SUCCEED, GTEST_SKIP, AssertionResult / AssertionSuccess / AssertionFailure,
//   InitGoogleTest (--gtest_filter=, --gtest_list_tests), RUN_ALL_TESTS, UnitTest::GetInstance()->listeners().Append,
//   EmptyTestEventListener::OnTestPartResult(TestPartResult).
// Output follows GoogleTest's "[ RUN ] / [ OK ] / [ FAILED ] / [ SKIPPED ]" lines so existing log readers keep working.
#ifndef CUDECOMP_B200_MINI_GTEST_H
#define CUDECOMP_B200_MINI_GTEST_H

#include <cstdio>
#include <cstring>
#include <functional>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <type_traits>
#include <vector>

namespace testing {

// ------------------------------------------------------------------------------------------ messages and results
class Message {
public:
  Message() = default;
  Message(const Message& o) { ss_ << o.str(); }
  template <typename T> Message& operator<<(const T& v) {
    ss_ << v;
    return *this;
  }
  Message& operator<<(std::ostream& (*manip)(std::ostream&)) {
    ss_ << manip;
    return *this;
  }
  Message& operator<<(bool b) {
    ss_ << (b ? "true" : "false");
    return *this;
  }
  std::string str() const { return ss_.str(); }

private:
  std::ostringstream ss_;
};

class AssertionResult {
public:
  explicit AssertionResult(bool ok) : ok_(ok) {}
  AssertionResult(const AssertionResult& o) : ok_(o.ok_), msg_(o.msg_) {}
  explicit operator bool() const { return ok_; }
  const char* message() const { return msg_.c_str(); }
  const char* failure_message() const { return msg_.c_str(); }
  template <typename T> AssertionResult& operator<<(const T& v) {
    std::ostringstream ss;
    ss << v;
    msg_ += ss.str();
    return *this;
  }

private:
  bool ok_;
  std::string msg_;
};
inline AssertionResult AssertionSuccess() { return AssertionResult(true); }
inline AssertionResult AssertionFailure() { return AssertionResult(false); }

class TestPartResult {
public:
  enum Type { kSuccess, kNonFatalFailure, kFatalFailure, kSkip };
  TestPartResult(Type t, const char* file, int line, std::string msg) : type_(t), file_(file ? file : ""), line_(line), msg_(std::move(msg)) {}
  Type type() const { return type_; }
  bool failed() const { return type_ == kNonFatalFailure || type_ == kFatalFailure; }
  bool skipped() const { return type_ == kSkip; }
  const char* file_name() const { return file_.c_str(); }
  int line_number() const { return line_; }
  const char* summary() const { return msg_.c_str(); }
  const char* message() const { return msg_.c_str(); }

private:
  Type type_;
  std::string file_;
  int line_;
  std::string msg_;
};

class TestEventListener {
public:
  virtual ~TestEventListener() = default;
  virtual void OnTestPartResult(const TestPartResult&) {}
};
class EmptyTestEventListener : public TestEventListener {};

class TestEventListeners {
public:
  void Append(TestEventListener* l) { listeners_.emplace_back(l); }
  void notify(const TestPartResult& r) {
    for (auto& l : listeners_) l->OnTestPartResult(r);
  }

private:
  std::vector<std::unique_ptr<TestEventListener>> listeners_;
};

// ------------------------------------------------------------------------------------------ tests and registry
class Test {
public:
  virtual ~Test() = default;
  virtual void SetUp() {}
  virtual void TearDown() {}
  virtual void TestBody() = 0;
  static bool HasFatalFailure();
  static bool HasFailure();
  static bool IsSkipped();
};

template <typename T> struct TestParamInfo {
  TestParamInfo(const T& p, size_t i) : param(p), index(i) {}
  T param;
  size_t index;
};

template <typename T> class WithParamInterface {
public:
  using ParamType = T;
  virtual ~WithParamInterface() = default;
  static const T& GetParam() { return *current_param(); }
  static const T*& current_param() {
    static const T* p = nullptr;
    return p;
  }
};
template <typename T> class TestWithParam : public Test, public WithParamInterface<T> {};

template <typename Container> struct ValuesInHolder {
  Container values;
};
template <typename Container> ValuesInHolder<Container> ValuesIn(const Container& c) { return ValuesInHolder<Container>{c}; }
template <typename T, size_t N> ValuesInHolder<std::vector<T>> ValuesIn(const T (&a)[N]) {
  return ValuesInHolder<std::vector<T>>{std::vector<T>(a, a + N)};
}
template <typename... Ts> auto Values(Ts... vs) {
  using T = typename std::common_type<Ts...>::type;
  return ValuesInHolder<std::vector<T>>{std::vector<T>{static_cast<T>(vs)...}};
}

struct RegisteredTest {
  std::string suite, name;
  const char* file;
  int line;
  std::function<void()> run; // constructs the fixture, SetUp, TestBody, TearDown
};

class UnitTest {
public:
  static UnitTest* GetInstance() {
    static UnitTest u;
    return &u;
  }
  TestEventListeners& listeners() { return listeners_; }
  std::vector<RegisteredTest>& tests() { return tests_; }
  // parameterised suites: bodies registered by TEST_P, expanded by INSTANTIATE_TEST_SUITE_P at RUN_ALL_TESTS time
  std::vector<std::function<void()>>& expanders() { return expanders_; }

  void report(TestPartResult::Type t, const char* file, int line, const std::string& msg) {
    TestPartResult r(t, file, line, msg);
    if (r.failed()) {
      cur_failed_ = true;
      if (t == TestPartResult::kFatalFailure) cur_fatal_ = true;
      std::printf("%s:%d: Failure\n%s\n", file ? file : "?", line, msg.c_str());
      std::fflush(stdout);
    } else if (r.skipped()) {
      cur_skipped_ = true;
      if (!msg.empty()) std::printf("%s:%d: Skipped\n%s\n", file ? file : "?", line, msg.c_str());
    }
    listeners_.notify(r);
  }
  bool cur_failed_ = false, cur_fatal_ = false, cur_skipped_ = false;
  std::string filter_ = "*";
  bool list_only_ = false;

  int Run();

private:
  TestEventListeners listeners_;
  std::vector<RegisteredTest> tests_;
  std::vector<std::function<void()>> expanders_;
};

inline bool Test::HasFatalFailure() { return UnitTest::GetInstance()->cur_fatal_; }
inline bool Test::HasFailure() { return UnitTest::GetInstance()->cur_failed_; }
inline bool Test::IsSkipped() { return UnitTest::GetInstance()->cur_skipped_; }

namespace internal {

// '*' and '?' wildcards; patterns separated by ':'; a '-' introduces the negative patterns (GoogleTest's filter syntax)
inline bool wildcardMatch(const char* p, const char* s) {
  if (*p == 0) return *s == 0;
  if (*p == '*') return wildcardMatch(p + 1, s) || (*s && wildcardMatch(p, s + 1));
  if (*s == 0) return false;
  return (*p == '?' || *p == *s) && wildcardMatch(p + 1, s + 1);
}
inline bool matchesAny(const std::string& patterns, const std::string& name) {
  size_t start = 0;
  while (start <= patterns.size()) {
    size_t end = patterns.find(':', start);
    if (end == std::string::npos) end = patterns.size();
    if (end > start && wildcardMatch(patterns.substr(start, end - start).c_str(), name.c_str())) return true;
    start = end + 1;
  }
  return false;
}
inline bool filterAccepts(const std::string& filter, const std::string& full) {
  const size_t dash = filter.find('-');
  const std::string pos = dash == std::string::npos ? filter : filter.substr(0, dash);
  const std::string neg = dash == std::string::npos ? "" : filter.substr(dash + 1);
  return matchesAny(pos.empty() ? "*" : pos, full) && !(neg.size() && matchesAny(neg, full));
}

template <typename Fixture> void runFixture() {
  Fixture f;
  Test& t = f; // fixtures usually declare SetUp / TearDown protected; virtual dispatch through the base reaches them
  t.SetUp();
  if (!Test::HasFatalFailure() && !Test::IsSkipped()) t.TestBody();
  t.TearDown();
}

struct Registrar {
  Registrar(const char* suite, const char* name, const char* file, int line, std::function<void()> run) {
    UnitTest::GetInstance()->tests().push_back({suite, name, file, line, std::move(run)});
  }
};

// bodies of a parameterised suite
template <typename Suite> struct ParamBodies {
  struct Body {
    std::string name;
    const char* file;
    int line;
    std::function<void()> run;
  };
  static std::vector<Body>& get() {
    static std::vector<Body> b;
    return b;
  }
};
template <typename Suite, typename Fixture> struct ParamRegistrar {
  ParamRegistrar(const char* name, const char* file, int line) {
    ParamBodies<Suite>::get().push_back({name, file, line, [] { runFixture<Fixture>(); }});
  }
};

template <typename T> std::string defaultParamName(const TestParamInfo<T>& info) { return std::to_string(info.index); }

template <typename Suite> struct Instantiator {
  using T = typename Suite::ParamType;
  template <typename Container, typename NameFn>
  Instantiator(const char* prefix, const char* suite, ValuesInHolder<Container> values, NameFn name_fn) {
    auto params = std::make_shared<std::vector<T>>(values.values.begin(), values.values.end());
    std::string pfx = prefix, sname = suite;
    UnitTest::GetInstance()->expanders().push_back([params, pfx, sname, name_fn] {
      for (auto& body : ParamBodies<Suite>::get()) {
        for (size_t i = 0; i < params->size(); ++i) {
          const std::string pname = name_fn(TestParamInfo<T>((*params)[i], i));
          const T* p = &(*params)[i];
          auto run = body.run;
          UnitTest::GetInstance()->tests().push_back({pfx + "/" + sname, body.name + "/" + pname, body.file, body.line, [params, p, run] {
                                                        WithParamInterface<T>::current_param() = p;
                                                        run();
                                                        WithParamInterface<T>::current_param() = nullptr;
                                                      }});
        }
      }
    });
  }
  template <typename Container> Instantiator(const char* prefix, const char* suite, ValuesInHolder<Container> values)
      : Instantiator(prefix, suite, values, &defaultParamName<T>) {}
};

// ---- assertion plumbing: `AssertHelper(...) = Message() << user text` runs when the statement ends
struct AssertHelper {
  AssertHelper(TestPartResult::Type t, const char* file, int line, std::string msg) : t_(t), file_(file), line_(line), msg_(std::move(msg)) {}
  void operator=(const Message& user) const {
    std::string m = msg_;
    const std::string u = user.str();
    if (!u.empty()) m += (m.empty() ? "" : "\n") + u;
    UnitTest::GetInstance()->report(t_, file_, line_, m);
  }
  TestPartResult::Type t_;
  const char* file_;
  int line_;
  std::string msg_;
};

template <typename T, typename = void> struct Printable : std::false_type {};
template <typename T> struct Printable<T, decltype(void(std::declval<std::ostream&>() << std::declval<const T&>()))> : std::true_type {};

template <typename T> std::string show(const T& v) {
  if constexpr (std::is_same<T, std::nullptr_t>::value) {
    return "(nullptr)";
  } else if constexpr (std::is_same<T, bool>::value) {
    return v ? "true" : "false";
  } else if constexpr (std::is_enum<T>::value) {
    return std::to_string(static_cast<long long>(v));
  } else if constexpr (std::is_pointer<T>::value && !std::is_same<T, const char*>::value && !std::is_same<T, char*>::value) {
    std::ostringstream ss;
    ss << static_cast<const void*>(v);
    return ss.str();
  } else if constexpr (Printable<T>::value) {
    std::ostringstream ss;
    ss << v;
    return ss.str();
  } else {
    return "<" + std::to_string(sizeof(T)) + "-byte object>";
  }
}

template <typename A, typename B, typename Op>
AssertionResult compare(const char* ea, const char* eb, const A& a, const B& b, const char* opname, Op op) {
  if (op(a, b)) return AssertionSuccess();
  return AssertionFailure() << "Expected: (" << ea << ") " << opname << " (" << eb << "), actual: " << show(a) << " vs " << show(b);
}
struct OpEq { template <typename A, typename B> bool operator()(const A& a, const B& b) const { return a == b; } };
struct OpNe { template <typename A, typename B> bool operator()(const A& a, const B& b) const { return a != b; } };
struct OpLt { template <typename A, typename B> bool operator()(const A& a, const B& b) const { return a < b; } };
struct OpLe { template <typename A, typename B> bool operator()(const A& a, const B& b) const { return a <= b; } };
struct OpGt { template <typename A, typename B> bool operator()(const A& a, const B& b) const { return a > b; } };
struct OpGe { template <typename A, typename B> bool operator()(const A& a, const B& b) const { return a >= b; } };

inline AssertionResult compareStr(const char* ea, const char* eb, const char* a, const char* b, bool want_equal) {
  const bool eq = (a == nullptr || b == nullptr) ? (a == b) : std::strcmp(a, b) == 0;
  if (eq == want_equal) return AssertionSuccess();
  return AssertionFailure() << "Expected " << (want_equal ? "equality" : "inequality") << " of these strings:\n  " << ea << " = \""
                            << (a ? a : "(null)") << "\"\n  " << eb << " = \"" << (b ? b : "(null)") << "\"";
}

inline AssertionResult boolResult(const AssertionResult& r, const char* expr, bool want) {
  if (static_cast<bool>(r) == want) return AssertionSuccess();
  AssertionResult f = AssertionFailure();
  f << "Value of: " << expr << "\n  Actual: " << (want ? "false" : "true");
  if (*r.message()) f << " (" << r.message() << ")";
  f << "\nExpected: " << (want ? "true" : "false");
  return f;
}
inline AssertionResult boolResult(bool v, const char* expr, bool want) { return boolResult(AssertionResult(v), expr, want); }
template <typename T> AssertionResult boolResult(const T& v, const char* expr, bool want) {
  return boolResult(AssertionResult(static_cast<bool>(v)), expr, want);
}

} // namespace internal

inline int UnitTest::Run() {
  for (auto& e : expanders_) e();
  expanders_.clear();
  std::vector<const RegisteredTest*> selected;
  for (auto& t : tests_)
    if (internal::filterAccepts(filter_, t.suite + "." + t.name)) selected.push_back(&t);
  if (list_only_) {
    for (auto* t : selected) std::printf("%s.%s\n", t->suite.c_str(), t->name.c_str());
    return 0;
  }
  std::printf("[==========] Running %zu tests.\n", selected.size());
  int failed = 0, skipped = 0, passed = 0;
  std::vector<std::string> failed_names;
  for (auto* t : selected) {
    const std::string full = t->suite + "." + t->name;
    cur_failed_ = cur_fatal_ = cur_skipped_ = false;
    std::printf("[ RUN      ] %s\n", full.c_str());
    std::fflush(stdout);
    try {
      t->run();
    } catch (const std::exception& e) {
      report(TestPartResult::kFatalFailure, t->file, t->line, std::string("C++ exception thrown in the test body: ") + e.what());
    } catch (...) {
      report(TestPartResult::kFatalFailure, t->file, t->line, "unknown C++ exception thrown in the test body");
    }
    if (cur_failed_) {
      ++failed;
      failed_names.push_back(full);
      std::printf("[  FAILED  ] %s\n", full.c_str());
    } else if (cur_skipped_) {
      ++skipped;
      std::printf("[  SKIPPED ] %s\n", full.c_str());
    } else {
      ++passed;
      std::printf("[       OK ] %s\n", full.c_str());
    }
    std::fflush(stdout);
  }
  std::printf("[==========] %zu tests ran.\n[  PASSED  ] %d tests.\n", selected.size(), passed);
  if (skipped) std::printf("[  SKIPPED ] %d tests.\n", skipped);
  if (failed) {
    std::printf("[  FAILED  ] %d tests, listed below:\n", failed);
    for (auto& n : failed_names) std::printf("[  FAILED  ] %s\n", n.c_str());
  }
  std::fflush(stdout);
  return failed ? 1 : 0;
}

inline void InitGoogleTest(int* argc, char** argv) {
  if (!argc || !argv) return;
  int out = 1;
  for (int i = 1; i < *argc; ++i) {
    const std::string a = argv[i];
    if (a.rfind("--gtest_filter=", 0) == 0) {
      UnitTest::GetInstance()->filter_ = a.substr(15);
    } else if (a == "--gtest_list_tests") {
      UnitTest::GetInstance()->list_only_ = true;
    } else if (a.rfind("--gtest_", 0) == 0) {
      // other GoogleTest flags (colour, brief, ...) are accepted and ignored
    } else {
      argv[out++] = argv[i];
    }
  }
  *argc = out;
}
inline void InitGoogleTest() {}

} // namespace testing

#define RUN_ALL_TESTS() (::testing::UnitTest::GetInstance()->Run())

// ------------------------------------------------------------------------------------------ test definition macros
#define GTEST_TEST_CLASS_NAME_(suite, name) suite##_##name##_Test

#define GTEST_TEST_(suite, name, parent)                                                                                \
  class GTEST_TEST_CLASS_NAME_(suite, name) : public parent {                                                           \
  public:                                                                                                               \
    void TestBody() override;                                                                                           \
  };                                                                                                                    \
  static ::testing::internal::Registrar gtest_registrar_##suite##_##name(                                               \
      #suite, #name, __FILE__, __LINE__, [] { ::testing::internal::runFixture<GTEST_TEST_CLASS_NAME_(suite, name)>(); });  \
  void GTEST_TEST_CLASS_NAME_(suite, name)::TestBody()

#define TEST(suite, name) GTEST_TEST_(suite, name, ::testing::Test)
#define TEST_F(fixture, name) GTEST_TEST_(fixture, name, fixture)

#define TEST_P(suite, name)                                                                                             \
  class GTEST_TEST_CLASS_NAME_(suite, name) : public suite {                                                            \
  public:                                                                                                               \
    void TestBody() override;                                                                                           \
  };                                                                                                                    \
  static ::testing::internal::ParamRegistrar<suite, GTEST_TEST_CLASS_NAME_(suite, name)> gtest_pregistrar_##suite##_##name( \
      #name, __FILE__, __LINE__);                                                                                       \
  void GTEST_TEST_CLASS_NAME_(suite, name)::TestBody()

#define INSTANTIATE_TEST_SUITE_P(prefix, suite, ...)                                                                    \
  static ::testing::internal::Instantiator<suite> gtest_instantiator_##prefix##_##suite(#prefix, #suite, __VA_ARGS__)
#define INSTANTIATE_TEST_CASE_P INSTANTIATE_TEST_SUITE_P

// ------------------------------------------------------------------------------------------ assertion macros
// The switch/else shape lets the macros sit inside unbraced if/else and lets the user stream a message after them.
#define GTEST_AMBIGUOUS_ELSE_BLOCKER_                                                                                   \
  switch (0)                                                                                                            \
  case 0:                                                                                                               \
  default:

#define GTEST_NONFATAL_(msg)                                                                                            \
  ::testing::internal::AssertHelper(::testing::TestPartResult::kNonFatalFailure, __FILE__, __LINE__, msg) = ::testing::Message()
#define GTEST_FATAL_(msg)                                                                                               \
  return ::testing::internal::AssertHelper(::testing::TestPartResult::kFatalFailure, __FILE__, __LINE__, msg) = ::testing::Message()

#define GTEST_ASSERT_(expression, on_failure)                                                                           \
  GTEST_AMBIGUOUS_ELSE_BLOCKER_                                                                                         \
  if (const ::testing::AssertionResult gtest_ar = (expression))                                                         \
    ;                                                                                                                   \
  else                                                                                                                  \
    on_failure(gtest_ar.failure_message())

#define GTEST_CMP_(a, b, opname, Op, on_failure)                                                                        \
  GTEST_ASSERT_(::testing::internal::compare(#a, #b, (a), (b), opname, ::testing::internal::Op()), on_failure)

#define EXPECT_EQ(a, b) GTEST_CMP_(a, b, "==", OpEq, GTEST_NONFATAL_)
#define EXPECT_NE(a, b) GTEST_CMP_(a, b, "!=", OpNe, GTEST_NONFATAL_)
#define EXPECT_LT(a, b) GTEST_CMP_(a, b, "<", OpLt, GTEST_NONFATAL_)
#define EXPECT_LE(a, b) GTEST_CMP_(a, b, "<=", OpLe, GTEST_NONFATAL_)
#define EXPECT_GT(a, b) GTEST_CMP_(a, b, ">", OpGt, GTEST_NONFATAL_)
#define EXPECT_GE(a, b) GTEST_CMP_(a, b, ">=", OpGe, GTEST_NONFATAL_)
#define ASSERT_EQ(a, b) GTEST_CMP_(a, b, "==", OpEq, GTEST_FATAL_)
#define ASSERT_NE(a, b) GTEST_CMP_(a, b, "!=", OpNe, GTEST_FATAL_)
#define ASSERT_LT(a, b) GTEST_CMP_(a, b, "<", OpLt, GTEST_FATAL_)
#define ASSERT_LE(a, b) GTEST_CMP_(a, b, "<=", OpLe, GTEST_FATAL_)
#define ASSERT_GT(a, b) GTEST_CMP_(a, b, ">", OpGt, GTEST_FATAL_)
#define ASSERT_GE(a, b) GTEST_CMP_(a, b, ">=", OpGe, GTEST_FATAL_)

#define EXPECT_TRUE(c) GTEST_ASSERT_(::testing::internal::boolResult((c), #c, true), GTEST_NONFATAL_)
#define EXPECT_FALSE(c) GTEST_ASSERT_(::testing::internal::boolResult((c), #c, false), GTEST_NONFATAL_)
#define ASSERT_TRUE(c) GTEST_ASSERT_(::testing::internal::boolResult((c), #c, true), GTEST_FATAL_)
#define ASSERT_FALSE(c) GTEST_ASSERT_(::testing::internal::boolResult((c), #c, false), GTEST_FATAL_)

#define EXPECT_STREQ(a, b) GTEST_ASSERT_(::testing::internal::compareStr(#a, #b, (a), (b), true), GTEST_NONFATAL_)
#define EXPECT_STRNE(a, b) GTEST_ASSERT_(::testing::internal::compareStr(#a, #b, (a), (b), false), GTEST_NONFATAL_)
#define ASSERT_STREQ(a, b) GTEST_ASSERT_(::testing::internal::compareStr(#a, #b, (a), (b), true), GTEST_FATAL_)
#define ASSERT_STRNE(a, b) GTEST_ASSERT_(::testing::internal::compareStr(#a, #b, (a), (b), false), GTEST_FATAL_)

#define GTEST_THROW_(statement, exception_type, on_failure)                                                             \
  GTEST_AMBIGUOUS_ELSE_BLOCKER_                                                                                         \
  if (const ::testing::AssertionResult gtest_ar = [&]() -> ::testing::AssertionResult {                                 \
        try {                                                                                                           \
          statement;                                                                                                    \
        } catch (const exception_type&) {                                                                               \
          return ::testing::AssertionSuccess();                                                                         \
        } catch (...) {                                                                                                 \
          return ::testing::AssertionFailure() << "Expected: " #statement " throws " #exception_type ", it throws a different type"; \
        }                                                                                                               \
        return ::testing::AssertionFailure() << "Expected: " #statement " throws " #exception_type ", it throws nothing"; \
      }())                                                                                                              \
    ;                                                                                                                   \
  else                                                                                                                  \
    on_failure(gtest_ar.failure_message())
#define EXPECT_THROW(statement, exception_type) GTEST_THROW_(statement, exception_type, GTEST_NONFATAL_)
#define ASSERT_THROW(statement, exception_type) GTEST_THROW_(statement, exception_type, GTEST_FATAL_)

#define GTEST_NO_THROW_(statement, on_failure)                                                                          \
  GTEST_AMBIGUOUS_ELSE_BLOCKER_                                                                                         \
  if (const ::testing::AssertionResult gtest_ar = [&]() -> ::testing::AssertionResult {                                 \
        try {                                                                                                           \
          statement;                                                                                                    \
        } catch (...) {                                                                                                 \
          return ::testing::AssertionFailure() << "Expected: " #statement " does not throw, it throws";                 \
        }                                                                                                               \
        return ::testing::AssertionSuccess();                                                                           \
      }())                                                                                                              \
    ;                                                                                                                   \
  else                                                                                                                  \
    on_failure(gtest_ar.failure_message())
#define EXPECT_NO_THROW(statement) GTEST_NO_THROW_(statement, GTEST_NONFATAL_)
#define ASSERT_NO_THROW(statement) GTEST_NO_THROW_(statement, GTEST_FATAL_)

#define ADD_FAILURE() GTEST_NONFATAL_("Failed")
#define FAIL() GTEST_FATAL_("Failed")
#define GTEST_FAIL() FAIL()
#define SUCCEED() ::testing::internal::AssertHelper(::testing::TestPartResult::kSuccess, __FILE__, __LINE__, "") = ::testing::Message()
#define GTEST_SUCCEED() SUCCEED()
#define GTEST_SKIP() return ::testing::internal::AssertHelper(::testing::TestPartResult::kSkip, __FILE__, __LINE__, "") = ::testing::Message()
#define SCOPED_TRACE(msg) (void)0

#endif
