// Test-infrastructure shim (oracle/Makefile, oracle/_ref only): lets UNMODIFIED reference sources compile without Boost.Filesystem.
#pragma once
#include <boost/filesystem.hpp>
