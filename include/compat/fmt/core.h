// <fmt/core.h> as the reference's suts_logger.h includes it: forwards to the fmt that spdlog bundles
// (the reference links spdlog and fmt as separate vcpkg packages; one fmt must serve both here).
#pragma once
#include <spdlog/fmt/fmt.h>
