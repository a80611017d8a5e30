// ORACLE — TEST INFRASTRUCTURE ONLY.
// Shadows source/util/RawUtil.h (479 lines of raw-sensor decoding, SURVEY.md §2: out of scope) for the oracle/_ref
// build: CvUtil.h:25 includes it only for rawToRgb(), which the depth path reaches for ".raw" inputs alone.
#pragma once
#include <opencv2/core.hpp>
#include "source/util/FilesystemUtil.h"
namespace fb360_dep {
inline cv::Mat rawToRgb(const filesystem::path&) { cv::shimUnsupported("rawToRgb (.raw input)"); }
}  // namespace fb360_dep
