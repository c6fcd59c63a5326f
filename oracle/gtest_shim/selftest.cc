// Self test of the GoogleTest stand-in (gtest/gtest.h): built and run by tests/test_gtest_shim.py on the CPU. It uses
// the features the reference's suites use and a few deliberate failures whose counts the Python side checks.
#include <stdexcept>
#include <string>
#include <vector>

#include <gtest/gtest.h>

namespace {
struct Case {
  int value;
  const char* label;
};
std::vector<Case> cases() { return {{1, "one"}, {2, "two"}, {3, "three"}}; }
std::string caseName(const testing::TestParamInfo<Case>& info) { return std::string(info.param.label) + "_" + std::to_string(info.index); }

testing::AssertionResult isEven(int v) {
  if (v % 2 == 0) return testing::AssertionSuccess();
  return testing::AssertionFailure() << v << " is odd";
}

int g_setups = 0, g_teardowns = 0, g_listener_failures = 0;

class Fixture : public ::testing::Test {
protected:
  void SetUp() override { ++g_setups; value_ = 41; }
  void TearDown() override { ++g_teardowns; }
  int value_ = 0;
};

class Listener : public ::testing::EmptyTestEventListener {
  void OnTestPartResult(const ::testing::TestPartResult& r) override {
    if (r.failed() && r.line_number() > 0 && r.file_name()[0] && r.summary()[0]) ++g_listener_failures;
  }
};

void helperThatAsserts(int v) { ASSERT_EQ(v, 7) << "helper got " << v; }
} // namespace

TEST(Plain, PassesWithEveryComparison) {
  EXPECT_EQ(2 + 2, 4);
  EXPECT_NE(1, 2);
  EXPECT_LT(1, 2);
  EXPECT_LE(2, 2);
  EXPECT_GT(3, 2) << "never shown";
  EXPECT_GE(3, 3);
  EXPECT_TRUE(isEven(4));
  EXPECT_FALSE(isEven(3));
  EXPECT_STREQ("abc", std::string("abc").c_str());
  EXPECT_STRNE("abc", "abd");
  void* p = nullptr;
  EXPECT_EQ(p, nullptr);
  ASSERT_NE(&p, nullptr);
  EXPECT_THROW(throw std::runtime_error("x"), std::runtime_error);
  EXPECT_NO_THROW((void)0);
  if (p == nullptr)
    EXPECT_TRUE(true);
  else
    EXPECT_TRUE(false);
  SUCCEED();
}

TEST(Plain, NonFatalFailuresContinue) {
  EXPECT_EQ(1, 2) << "first";
  EXPECT_TRUE(isEven(5)) << "second";
  EXPECT_STREQ("a", "b");
  EXPECT_THROW((void)0, std::runtime_error);
  std::printf("REACHED_AFTER_NONFATAL\n");
}

TEST(Plain, FatalFailureReturns) {
  ASSERT_TRUE(false) << "stops here";
  std::printf("NOT_REACHED\n");
}

TEST(Plain, FatalInHelperOnlyLeavesTheHelper) {
  helperThatAsserts(8);
  EXPECT_TRUE(::testing::Test::HasFatalFailure());
  std::printf("REACHED_AFTER_HELPER\n");
}

TEST(Plain, Skips) {
  GTEST_SKIP() << "not today";
  std::printf("NOT_REACHED\n");
}

TEST(Plain, ExceptionIsAFailure) { throw std::runtime_error("boom"); }

TEST_F(Fixture, SeesSetUp) { EXPECT_EQ(value_, 41); }
TEST_F(Fixture, SeesSetUpAgain) { ASSERT_EQ(value_, 41); }

class Param : public ::testing::TestWithParam<Case> {};
TEST_P(Param, ValueMatchesLabel) {
  const Case& c = GetParam();
  switch (c.value) {
  case 1: EXPECT_STREQ(c.label, "one"); break;
  case 2: EXPECT_STREQ(c.label, "two"); break;
  case 3: EXPECT_STREQ(c.label, "three"); break;
  default: FAIL() << "unexpected value " << c.value;
  }
}
TEST_P(Param, OddValuesFail) { EXPECT_TRUE(isEven(GetParam().value)); }
INSTANTIATE_TEST_SUITE_P(First, Param, ::testing::ValuesIn(cases()), caseName);
INSTANTIATE_TEST_SUITE_P(Second, Param, ::testing::ValuesIn(cases()));

int main(int argc, char** argv) {
  ::testing::InitGoogleTest(&argc, argv);
  ::testing::UnitTest::GetInstance()->listeners().Append(new Listener);
  const int rc = RUN_ALL_TESTS();
  std::printf("SETUPS=%d TEARDOWNS=%d LISTENER_FAILURES=%d RC=%d ARGC=%d\n", g_setups, g_teardowns, g_listener_failures, rc, argc);
  return 0;
}
