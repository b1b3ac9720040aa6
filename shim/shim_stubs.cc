// Link-time stubs for two reference subsystems that are out of scope and not built into the shim
// harness: the Python runtime factory and the YAML codec (flowgraphs are built through the C++ API here).
#include <memory>
#include <string>

#include "jetstream/parser.hh"
#include "jetstream/runtime.hh"

namespace Jetstream {

std::shared_ptr<Runtime::Impl> PythonRuntimeFactory() { return nullptr; }

Result Parser::YamlEncode(const Map&, std::string&) { return Result::ERROR; }
Result Parser::YamlDecode(const std::string&, Map&) { return Result::ERROR; }

}  // namespace Jetstream
