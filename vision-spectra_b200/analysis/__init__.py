"""Delta-alpha / publication summary straight from the gathered per-epoch records."""

from .summary import (  # noqa: F401
    SCENARIO_METADATA,
    RunHistory,
    ScenarioMetrics,
    perform_statistical_tests,
    run_history_from_epochs,
    scenario_metrics,
    students_t_test,
)
