// Host-only part of the C ABI (layout queries).  Included inside an extern "C" block by
// muav_kernels.cu (the product library) and by tests/hostcheck/hostcheck.cpp (CPU build of the
// same simulation core, used by the `-m "not gpu"` tests only).
const char* muav_version(void) { return "muav_b200 0.1 (sm_100a)"; }
size_t muav_config_size(void) { return sizeof(muav_config); }
size_t muav_record_bytes(const muav_config* cfg) { return (size_t)make_layout(*cfg).record_bytes; }
size_t muav_scratch_bytes(const muav_config* cfg) { return (size_t)make_layout(*cfg).scratch_bytes; }
size_t muav_hot_bytes(const muav_config* cfg) { return (size_t)make_layout(*cfg).hot_bytes; }

int muav_num_fields(void) {
  int n = 0;
#define X(name, type, count) ++n;
  MUAV_FIELDS(X, X, X, X)
#undef X
  return n;
}

int muav_field_info(const muav_config* cfg, int idx, const char** name, int64_t* offset, int64_t* count, int32_t* elem_size) {
  Layout L = make_layout(*cfg);
  Dims D = L.D;
  (void)D;
  int i = 0;
#define X(fname, type, cnt)             \
  if (i == idx) {                       \
    *name = #fname;                     \
    *offset = L.o_##fname;              \
    *count = (int64_t)(cnt);            \
    *elem_size = (int32_t)sizeof(type); \
    return 0;                           \
  }                                     \
  ++i;
  MUAV_FIELDS(X, X, X, X)
#undef X
  return -22;
}

int muav_header_index(const char* name) {
  int i = 0;
#define X(n)                                   \
  if (strcmp(name, #n) == 0) return i;         \
  ++i;
  MUAV_HI_LIST(X)
#undef X
  i = 0;
#define X(n)                                   \
  if (strcmp(name, #n) == 0) return i;         \
  ++i;
  MUAV_HF_LIST(X)
#undef X
  return -1;
}

static int check_cfg(const muav_config* c) {
  if (!c) return -22;
  if (c->n_agents < 1 || c->n_agents > MUAV_MAX_AGENTS) return -22;
  if (c->task_cap < 1 || c->task_cap > MUAV_MAX_TASK_CAP) return -22;
  if (c->id_cap > MUAV_MAX_ID_CAP) return -22;
  if (c->queue_cap < 1 || c->queue_cap > MUAV_MAX_QUEUE) return -22;
  if (c->n_groups < 0 || c->n_groups > MUAV_MAX_GROUPS) return -22;
  if (c->n_threats < 0 || c->event_cap < 1 || c->n_obstacles < 0) return -22;
  return 0;
}


// metric order of muav_metrics (calculate_metrics, DroneEnv.py:1286-1319)
static const char* const kMetricNames[MUAV_N_METRICS] = {
    "F_time", "F_distance", "F_quality", "F_Reward", "S_WPS", "S_ESC", "Losses", "Kills", "makespan", "total_distance",
    "n_reallocations", "n_task_switches", "n_arrivals", "n_tasks_final", "n_reached", "n_missed_windows", "n_on_time",
    "n_windowed_tasks", "on_time_rate", "reserve_idle_fraction", "escort_coverage_rate", "protected_rec_completed",
    "recon_losses", "escort_losses", "threats_intercepted", "mutual_support_engagements", "protection_breaches",
    "escort_requests", "escort_completed", "escort_failed"};


const char* muav_metric_name(int idx) { return (idx >= 0 && idx < MUAV_N_METRICS) ? kMetricNames[idx] : nullptr; }
