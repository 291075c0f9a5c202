"""The reference's OWN unit tests (dpose_core/test/*.cpp), compiled unmodified (tests/cpp/reference_gtests.py):
on the reference's sources (CPU, shows the stand-in headers carry them) and on this repository's drop-in headers +
libdpose_b200.so (GPU, the drop-in proof).  Skipped where neither /root/reference nor the prebuilt binaries exist."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "cpp"))
import reference_gtests as rg  # noqa: E402


@pytest.fixture(scope="module")
def on_reference():
    return rg.build_on_reference()


@pytest.fixture(scope="module")
def on_b200():
    return rg.build_on_b200()


@pytest.mark.parametrize("name", sorted(rg.SUITES))
def test_reference_tests_pass_on_reference_sources(on_reference, name):
    if name not in on_reference:
        pytest.skip("oracle/_ref/tests not built (no /root/reference here)")
    rc, ran, failed, out = rg.run(on_reference[name])
    assert rc == 0 and failed == 0, out
    assert ran == rg.SUITES[name], (ran, out)


def test_gtest_standin_reports_failures(tmp_path):
    """the stand-in GoogleTest must fail when a test fails (EXPECT and ASSERT, plain, fixture and parameterised)"""
    src = tmp_path / "t.cpp"
    src.write_text('''
#include <gtest/gtest.h>
struct fx : public testing::TestWithParam<int> {};
INSTANTIATE_TEST_SUITE_P(/**/, fx, testing::Range(0, 4));
TEST_P(fx, odd_fails) { EXPECT_EQ(GetParam() % 2, 0) << "param " << GetParam(); }
TEST(plain, passes) { EXPECT_TRUE(true); ASSERT_EQ(1, 1); }
TEST(plain, assert_leaves) { ASSERT_TRUE(false); std::abort(); }
TEST(plain, no_throw_detected) { EXPECT_ANY_THROW((void)0); }
''')
    exe = tmp_path / "t"
    subprocess.run([rg.CXX, "-std=c++17", "-I", rg.SHIM, str(src), os.path.join(rg.SHIM, "gtest", "gtest_main.cpp"),
                    "-o", str(exe)], check=True)
    rc, ran, failed, out = rg.run(str(exe))
    assert rc == 1 and ran == 7 and failed == 4, out
    assert "param 1" in out and "param 3" in out


def test_drop_in_headers_compile_the_reference_tests(on_b200):
    """compile-only on CPU: the reference's test sources build against include/dpose_core/*.hpp"""
    if not on_b200:
        pytest.skip("no /root/reference here and no prebuilt binaries")
    assert sorted(on_b200) == sorted(rg.SUITES)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(rg.SUITES))
def test_reference_tests_pass_on_b200(on_b200, name):
    if name not in on_b200:
        pytest.skip("tests/cpp/_build/b200_ref_* did not travel")
    rc, ran, failed, out = rg.run(on_b200[name])
    assert rc == 0 and failed == 0, out
    assert ran == rg.SUITES[name], (ran, out)
