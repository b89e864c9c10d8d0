// speck_b200/host/RunConfig.cpp -- argv handling of runspECK (reference source/RunConfig.cpp:8-23):
// argv[1] = matrix path, argv[2] = ini file; the ini key InputFile overrides argv[1].
#include "Config.h"
#include "RunConfig.h"

RunConfig::RunConfig(int argc, char *argv[])
{
    if (argc < 2) throw "No file path set\n";
    filePath = argv[1];
    if (argc > 2) Config::init(std::string(argv[2]));
    else Config::init();
    filePath = Config::getString(Config::InputFile, filePath);
}
