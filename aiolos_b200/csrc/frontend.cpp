#include "problem.h"
namespace aio {
bool problem_from_files(const std::string&, const std::string&, const std::string&,
                        const std::vector<std::pair<std::string, std::string>>&, Problem&, std::string& err) {
    err = "unsupported configuration: front end not built yet";
    return false;
}
}
