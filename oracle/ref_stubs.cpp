// oracle/ref_stubs.cpp — TEST INFRASTRUCTURE. Link-time no-ops for the reference's optional
// Chrome-trace profiler (llvm TimeProfiler; off unless `tt.enable`, A/projects/spades/main.cpp:112-117).
// With the profiler instance null, LLVM's own implementations of these entry points do nothing
// (assembler/ext/src/llvm/TimeProfiler.cpp:315-334), so stubbing them leaves the computing path
// untouched while avoiding ~60 llvm-support translation units.
#include <llvm/Support/TimeProfiler.h>
namespace llvm {
TimeTraceProfiler *getTimeTraceProfilerInstance() { return nullptr; }
void timeTraceProfilerBegin(StringRef, StringRef) {}
void timeTraceProfilerEnd() {}
}  // namespace llvm
