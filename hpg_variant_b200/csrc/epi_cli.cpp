// epi_cli.cpp -- `hpg-var-gwas-b200 epi ...`: the one-line dispatch of src/gwas/main_gwas.c:23-98 that this path needs
// (log file, configuration file discovery in --config dir, cwd, ~/.hpg-variant, /etc/hpg-variant; then epistasis()).
#include "../../include/hpgv_epi_compat.h"

#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <string>
#include <vector>

static std::string find_config(int argc, char *argv[]) {
    std::vector<std::string> dirs;
    for (int a = 1; a < argc; a++) {
        std::string arg = argv[a];
        if ((arg == "--config" || arg == "-c") && a + 1 < argc) dirs.push_back(argv[a + 1]);
        else if (arg.rfind("--config=", 0) == 0) dirs.push_back(arg.substr(9));
    }
    char cwd[4096];
    if (getcwd(cwd, sizeof cwd)) dirs.push_back(cwd);
    if (const char *home = getenv("HOME")) dirs.push_back(std::string(home) + "/.hpg-variant");
    dirs.push_back("/etc/hpg-variant");
    struct stat sb;
    for (const std::string &d : dirs) {
        if (!stat(d.c_str(), &sb) && S_ISREG(sb.st_mode)) return d;           // --config may name the file itself
        const std::string f = d + "/hpg-variant.conf";
        if (!stat(f.c_str(), &sb) && S_ISREG(sb.st_mode)) return f;
    }
    return "";
}

int main(int argc, char *argv[]) {
    if (argc < 2 || !strcmp(argv[1], "-h") || !strcmp(argv[1], "--help")) {
        printf("Usage: %s epi [options]   (B200 build of `hpg-var-gwas epi`; `%s epi --help` lists the options)\n", argv[0], argv[0]);
        return 0;
    }
    if (strcmp(argv[1], "epi") != 0) {
        fprintf(stderr, "The requested genome-wide analysis tool does not exist! (%s)\n", argv[1]);
        return 1;
    }
    hpgv_epi_host_open_log("hpg-var-gwas.log");
    const std::string config = find_config(argc, argv);
    const int exit_code = epistasis(argc - 1, argv + 1, config.empty() ? NULL : config.c_str());
    if (exit_code > 0) fprintf(stderr, "Tool %s terminated with failure (exit code = %d)\n", argv[1], exit_code);
    hpgv_epi_host_open_log(NULL);
    return exit_code;
}
