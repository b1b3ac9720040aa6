// TEST INFRASTRUCTURE (oracle/_ref build only). Two link-time stubs for reference
// subsystems that are out of scope and whose third-party deps (CPython, rapidyaml)
// are absent: the Python runtime factory (src/runtime/runtime.cc:10,29) and the YAML
// codec (include/jetstream/parser.hh:209-210). Neither is on the DSP compute path.
#include <memory>
#include <string>

#include "jetstream/parser.hh"
#include "jetstream/runtime.hh"

namespace Jetstream {

std::shared_ptr<Runtime::Impl> PythonRuntimeFactory() { return nullptr; }

Result Parser::YamlEncode(const Map&, std::string&) { return Result::ERROR; }
Result Parser::YamlDecode(const std::string&, Map&) { return Result::ERROR; }

}  // namespace Jetstream
