// Test-infrastructure shim: lets the reference's searchTree.cc / normals.cc compile
// without Boost.  Only used by oracle/Makefile when building oracle/_ref.
#pragma once
#include <mutex>
namespace boost { using mutex = std::mutex; }
