// ORACLE — TEST INFRASTRUCTURE ONLY.  The library sources of the depth path include gflags but define no flags.
#pragma once
