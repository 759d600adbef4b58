// Test-infrastructure shim (oracle/Makefile, oracle/_ref only): lets UNMODIFIED reference sources compile without the scan server (shared-memory scans, include/scanserver: out of scope); shadows include/slam6d/managedScan.h.
#pragma once
// shim: scan-server (shared-memory) scans are out of scope; only what scan.cc names
#include "slam6d/scan.h"
class ManagedScan : public Scan {
 public:
  static void openDirectory(const std::string&, IOType, int, int) {}
  static void closeDirectory() {}
  void setShowReductionParameter(double, int) {}
};
