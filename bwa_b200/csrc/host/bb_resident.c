/* bb_resident.c -- an index that stays on the GPU between runs: `bwa-b200 shm` and the lookup `bwa-b200 mem` does first.
 *
 * The reference parks an index in POSIX shared memory (`bwa shm idxbase`, bwashm.c:16-122; `bwa_idx_load_from_shm`, bwa.c:337-401)
 * so that later `bwa mem` runs skip reading and unpacking the files.  On the GPU the expensive part is placing the index in HBM
 * (3 Gbp: read 5 GB of files, upload, re-pack the Occ table, derive the dense suffix-array sample, build the short-string table:
 * tens of seconds), and device memory lives only as long as its process.  So:
 *
 *   bwa-b200 shm idxbase     starts a keeper process: it loads the index the usual way, lets the device stage export it (CUDA IPC
 *                            handles, bwag_ctx_export) and sleeps; the command returns once the index is resident
 *   bwa-b200 mem idxbase ... finds the keeper's descriptor, loads only .ann/.amb/.pac from disk and attaches to the keeper's device
 *                            memory (bwag_ctx_import): no .bwt/.sa read, nothing uploaded, the start-up self-check still runs
 *   bwa-b200 shm -l          lists the resident indexes;   bwa-b200 shm -d   ends the keepers (and frees the HBM)
 *
 * Descriptors live in $BWA_B200_SHM_DIR (default /dev/shm), one per index, named by user id and a hash of the index's real path.
 */
#define _GNU_SOURCE
#include <dirent.h>
#include <fcntl.h>
#include <errno.h>
#include <signal.h>
#include <sys/stat.h>
#include <sys/wait.h>
#include <unistd.h>
#include <limits.h>
#include "bb_host.h"
#include "bwa_b200_dev.h"

static const char *share_dir(void) { const char *d = getenv("BWA_B200_SHM_DIR"); return d && *d ? d : "/dev/shm"; }

static void descriptor_path(const char *prefix, char *out, size_t cap)
{
	char real[PATH_MAX];
	uint64_t h = 1469598103934665603ULL;   /* FNV-1a of the real path of the .ann file's directory entry */
	const char *p;
	char ann[PATH_MAX];
	snprintf(ann, sizeof(ann), "%s.ann", prefix);
	if (!realpath(ann, real)) snprintf(real, sizeof(real), "%s", ann);
	for (p = real; *p; ++p) { h ^= (unsigned char)*p; h *= 1099511628211ULL; }
	snprintf(out, cap, "%s/bwa_b200.%u.%016llx.resident", share_dir(), (unsigned)getuid(), (unsigned long long)h);
}

/* `bwa-b200 mem` side: an index object over the keeper's device memory, or NULL if there is no (live) keeper for this prefix */
bwaidx_t *bb_idx_from_resident(const char *prefix)
{
	char path[PATH_MAX];
	struct stat st;
	bwaidx_t *idx;
	bwag_ctx_t *ctx;
	const char *e = getenv("BWA_B200_RESIDENT");
	if (e && atoi(e) == 0) return 0;
	descriptor_path(prefix, path, sizeof(path));
	if (stat(path, &st) != 0) return 0;
	idx = bwa_idx_load(prefix, BWA_IDX_BNS | BWA_IDX_PAC);
	if (!idx) return 0;
	ctx = bwag_ctx_import(path, idx->bns->l_pac);
	if (!ctx) {
		if (bwa_verbose >= 2) fprintf(stderr, "[W::%s] not using the resident index: %s\n", __func__, bwag_last_error());
		bwa_idx_destroy(idx);
		return 0;
	}
	idx->bwt = bb_calloc(1, sizeof(bwt_t));   /* no FM-index in host memory: the object only keys the device copy */
	bb_device_adopt2(idx->bwt, idx->bns, idx->pac, ctx);
	if (bwa_verbose >= 3) fprintf(stderr, "[M::%s] using the index resident on the GPU (%s)\n", __func__, path);
	return idx;
}

static volatile sig_atomic_t g_keeper_stop;
static void on_term(int sig) { (void)sig; g_keeper_stop = 1; }

static int keeper(const char *prefix, const char *path, int ready_fd)
{
	bwaidx_t *idx = bwa_idx_load(prefix, BWA_IDX_ALL);
	bwag_ctx_t *ctx;
	char ok = 1;
	struct sigaction sa;
	if (!idx) return 1;
	ctx = bb_device_attach(idx->bwt, idx->bns, idx->pac);   /* upload, dense sample, short-string table, self-check */
	if (bwag_ctx_export(ctx, path) != 0) { fprintf(stderr, "[E::%s] %s\n", "bwa_shm", bwag_last_error()); return 1; }
	/* the FM-index in host memory has served its purpose; the contig table and the packed text stay (cheap, and they keep bwa_idx_destroy simple) */
	free(idx->bwt->bwt); idx->bwt->bwt = 0; free(idx->bwt->sa); idx->bwt->sa = 0;
	memset(&sa, 0, sizeof(sa)); sa.sa_handler = on_term;
	sigaction(SIGTERM, &sa, 0); sigaction(SIGINT, &sa, 0); sigaction(SIGHUP, &sa, 0);
	{   /* from here on a daemon: let go of the starter's terminal and pipes (whoever waits for them to close would wait for ever) */
		int nul = open("/dev/null", O_RDWR);
		fflush(stdout); fflush(stderr);
		if (nul >= 0) { dup2(nul, 0); dup2(nul, 1); dup2(nul, 2); if (nul > 2) close(nul); }
	}
	if (write(ready_fd, &ok, 1) != 1) { /* the starter is gone: stay resident anyway */ }
	close(ready_fd);
	while (!g_keeper_stop) pause();
	bwag_ctx_unexport(path);
	bwa_idx_destroy(idx);
	return 0;
}

typedef struct { char path[PATH_MAX]; int pid; } resident_t;
static int list_residents(resident_t **out)   /* descriptors of this user in the share directory */
{
	DIR *d = opendir(share_dir());
	struct dirent *e;
	char mine[64];
	int n = 0, m = 0;
	*out = 0;
	if (!d) return 0;
	snprintf(mine, sizeof(mine), "bwa_b200.%u.", (unsigned)getuid());
	while ((e = readdir(d)) != 0) {
		size_t l = strlen(e->d_name);
		FILE *fp;
		int32_t hdr[5];
		if (strncmp(e->d_name, mine, strlen(mine)) != 0 || l < 9 || strcmp(e->d_name + l - 9, ".resident") != 0) continue;
		if (n == m) { m = m ? m << 1 : 8; *out = bb_realloc(*out, (size_t)m * sizeof(resident_t)); }
		snprintf((*out)[n].path, PATH_MAX, "%s/%s", share_dir(), e->d_name);
		(*out)[n].pid = -1;
		if ((fp = fopen((*out)[n].path, "rb")) != 0) {   /* magic[8], version, device, pid, ... (bwag_api.cu: ShareFile) */
			if (fread(hdr, 4, 5, fp) == 5) (*out)[n].pid = hdr[4];
			fclose(fp);
		}
		++n;
	}
	closedir(d);
	return n;
}

int bb_shm_main(int argc, char *argv[])
{
	int c, to_list = 0, to_drop = 0;
	while ((c = getopt(argc, argv, "ldf:")) >= 0) {
		if (c == 'l') to_list = 1;
		else if (c == 'd') to_drop = 1;
		else if (c == 'f') { /* the reference's temporary-file option: nothing to stage through a file here */ }
		else return 1;
	}
	if (optind == argc && !to_list && !to_drop) {
		fprintf(stderr, "\nUsage: bwa-b200 shm [-d|-l] [idxbase]\n\nOptions: -d       end the processes that keep indexes resident on the GPU\n         -l       list the resident indexes\n\n");
		return 1;
	}
	if (optind < argc && (to_list || to_drop)) { fprintf(stderr, "[E::%s] open -l or -d cannot be used when 'idxbase' is present\n", __func__); return 1; }
	if (optind < argc) {
		char path[PATH_MAX], ok = 0;
		struct stat st;
		int pfd[2];
		pid_t pid;
		descriptor_path(argv[optind], path, sizeof(path));
		if (stat(path, &st) == 0) {
			resident_t *r; int i, n = list_residents(&r), alive = 0;
			for (i = 0; i < n; ++i) if (strcmp(r[i].path, path) == 0 && r[i].pid > 0 && (kill(r[i].pid, 0) == 0 || errno == EPERM)) alive = 1;
			free(r);
			if (alive) { fprintf(stderr, "[M::%s] index '%s' is already resident\n", __func__, argv[optind]); return 0; }
			bwag_ctx_unexport(path);   /* a keeper that died without cleaning up */
		}
		if (pipe(pfd) != 0) bb_fatal("bwa_shm", "pipe: %s", strerror(errno));
		pid = fork();   /* before anything touches CUDA in this process */
		if (pid < 0) bb_fatal("bwa_shm", "fork: %s", strerror(errno));
		if (pid == 0) {
			close(pfd[0]);
			setsid();
			_exit(keeper(argv[optind], path, pfd[1]));
		}
		close(pfd[1]);
		if (read(pfd[0], &ok, 1) != 1 || !ok) {   /* the keeper exited (its message is on stderr) */
			int status;
			waitpid(pid, &status, 0);
			fprintf(stderr, "[E::%s] failed to make '%s' resident\n", __func__, argv[optind]);
			return 1;
		}
		if (bwa_verbose >= 3) fprintf(stderr, "[M::%s] index '%s' is resident on the GPU (keeper pid %d)\n", __func__, argv[optind], (int)pid);
		return 0;
	}
	{
		resident_t *r; int i, n = list_residents(&r);
		for (i = 0; i < n; ++i) {
			const int alive = r[i].pid > 0 && (kill(r[i].pid, 0) == 0 || errno == EPERM);
			if (to_list) printf("%s\t%d\t%s\n", r[i].path, r[i].pid, alive ? "resident" : "stale");
			if (to_drop) {
				if (alive) {
					int k;
					kill(r[i].pid, SIGTERM);
					for (k = 0; k < 2000 && (kill(r[i].pid, 0) == 0 || errno == EPERM); ++k) usleep(5000);   /* the keeper removes its descriptor itself */
				}
				bwag_ctx_unexport(r[i].path);
			}
		}
		free(r);
	}
	return 0;
}
